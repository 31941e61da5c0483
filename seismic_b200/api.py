"""`seismic`-compatible Python surface (reference src/pylib/mod.rs, src/pylib/dataset.rs, src/pylib/dotvbyte.rs).

Same class names, method names, argument meaning and return shapes as the reference's PyO3 module, so a script
written against `import seismic` runs with `import seismic_b200 as seismic`:

    SeismicIndex / SeismicIndexLV            string token + string doc-id indexes  (src/pylib/mod.rs:46-661)
    SeismicIndexRaw / SeismicIndexRawLV      integer-component indexes over .bin    (src/pylib/mod.rs:663-1151)
    SeismicIndexDotVByte                     compressed forward index               (src/pylib/dotvbyte.rs)
    SeismicDataset / SeismicDatasetLV        growable dataset + brute-force search  (src/pylib/dataset.rs)
    get_seismic_string()                     "U30"                                  (src/pylib/mod.rs:24-25,41-44)

The index is BUILT on the CPU (C++ restatement of the reference build) and SEARCHED on the GPU: search /
batch_search marshal the queries into one CSR batch and make ONE call into the C ABI (sgpu_batch_search), i.e. the
reference's rayon loop over InvertedIndexBase::search is replaced by the persistent CUDA kernel.  There is no CPU
search path in this package: without a CUDA device search raises.
"""
from __future__ import annotations

import gzip
import io
import json
import os
import tarfile
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N
from .core import Dataset, GpuIndex, HostIndex, csr_from_lists, make_config

SEISMIC_STRING = "U30"
MAX_TOKEN_LEN = 30


def get_seismic_string() -> str:
    return SEISMIC_STRING


# ------------------------------------------------------------------------------------------ JSONL ingestion
def _open_collection(path: str) -> io.TextIOBase:
    """.jsonl or .tar.gz whose first member is the jsonl (reference src/inverted_index_wrapper.rs:526-596)."""
    if path.endswith(".jsonl"):
        return open(path, "r", encoding="utf-8")
    if path.endswith(".tar.gz"):
        tf = tarfile.open(path, "r:gz")
        member = tf.next()
        return io.TextIOWrapper(tf.extractfile(member), encoding="utf-8")
    raise OSError("Unsupported file type. Supported files: .jsonl, .tar.gz")


def _iter_jsonl(path: str):
    """records {id, vector:{token:score}, content?} (reference src/json_utils.rs:17-61); ids may be int or str."""
    with _open_collection(path) as f:
        for line in f:
            line = line.strip()
            if line:
                rec = json.loads(line)
                yield str(rec["id"]), rec["vector"], rec.get("content")


def _read_collection(path: str, token_to_id: Optional[Dict[str, int]], load_content: bool, max_tokens: int):
    """Two passes like the reference: token map in first-seen order, then vectors sorted by component."""
    if token_to_id is None:
        token_to_id = {}
        for _, vec, _ in _iter_jsonl(path):
            for tok in vec:
                if tok not in token_to_id:
                    token_to_id[tok] = len(token_to_id)
        if len(token_to_id) >= max_tokens:
            raise ValueError("The number of different tokens exceeds %d." % max_tokens)
    doc_ids: List[str] = []
    contents: Optional[List[Optional[str]]] = [] if load_content else None
    comps, vals = [], []
    for doc_id, vec, content in _iter_jsonl(path):
        doc_ids.append(doc_id)
        if contents is not None:
            contents.append(content)
        c = np.fromiter((token_to_id[t] for t in vec), dtype=np.uint32, count=len(vec))
        v = np.fromiter(vec.values(), dtype=np.float32, count=len(vec))
        order = np.argsort(c, kind="stable")
        comps.append(c[order])
        vals.append(v[order])
    return token_to_id, doc_ids, contents, comps, vals


# ------------------------------------------------------------------------------------------ shared machinery
class _IndexBase:
    """Host index + lazily created HBM image + the one-call batched search."""
    _COMP_BITS = 16
    _VALUE_KIND = N.VAL_F16

    def __init__(self, host: HostIndex, device: int = 0):
        self._host = host
        self._device = device
        self._gpu: Optional[GpuIndex] = None

    # -- getters (reference src/pylib/mod.rs:76-127)
    @property
    def dim(self) -> int:
        return self._host.dim

    @property
    def len(self) -> int:
        return self._host.len

    @property
    def nnz(self) -> int:
        return self._host.nnz

    @property
    def knn_len(self) -> int:
        """Neighbours per document of the attached kNN graph, 0 if none (src/inverted_index.rs:421-426)."""
        return 0 if self._host.knn is None else int(self._host.knn.shape[1])

    @property
    def is_empty(self) -> bool:
        return self._host.len == 0

    def __len__(self) -> int:
        return self._host.len

    def get(self, id: int) -> Tuple[List[int], List[float]]:
        c, v = self._host.get_doc(id)
        return c.tolist(), v.tolist()

    def get_doc_ids_in_postings(self, list_id: int) -> List[int]:
        if not 0 <= list_id < self.dim:
            raise ValueError("Invalid list_id: %d" % list_id)
        a = self._host.arrays()
        lo, hi = int(a["list_post_start"][list_id]), int(a["list_post_start"][list_id + 1])
        starts = (a["postings"][lo:hi] >> np.uint64(16)).astype(np.int64)
        return (np.searchsorted(a["fwd_offsets"].astype(np.int64), starts, side="right") - 1).tolist()

    def print_space_usage_byte(self) -> None:
        """Same lines as the reference (src/inverted_index.rs:103-149); the harness greps `\\tTotal: (\\d+) Bytes`."""
        u = self._host.space_usage()
        post = u["packed_postings"] + u["block_offsets"] + u["summaries"]
        pct = lambda x: 100.0 * x / post if post else 0.0  # noqa: E731
        print("Space Usage:")
        print("\tForward Index: %d Bytes" % u["forward"])
        print("\tPosting Lists: %d Bytes" % post)
        print("\t  ├─ packed_postings: %d Bytes (%.2f%%)" % (u["packed_postings"], pct(u["packed_postings"])))
        print("\t  ├─ block_offsets: %d Bytes (%.2f%%)" % (u["block_offsets"], pct(u["block_offsets"])))
        print("\t  └─ summaries: %d Bytes (%.2f%%)" % (u["summaries"], pct(u["summaries"])))
        print("\tKnn: %d Bytes" % (0 if self._host.knn is None else self._host.knn.nbytes))
        print("\tTotal: %d Bytes" % u["total"])

    # -- kNN graph (reference Knn, src/inverted_index.rs:430-593; Python surface src/pylib/mod.rs:224-291)
    _KNN_QUERY_CUT, _KNN_HEAP_FACTOR = 10, 0.7  # src/inverted_index.rs:456-457
    _KNN_MAGIC = b"SB2KNN01"

    def _attach_knn(self, neighbours: Optional[np.ndarray]) -> None:
        self._host.set_knn(neighbours)
        if self._gpu is not None:
            self._gpu.set_knn(neighbours)

    def build_knn(self, nknn: int, batch_docs: int = 200_000) -> None:
        """Knn::new (src/inverted_index.rs:448-500): every document searches the index with its own vector
        (k = nknn + 1, query_cut 10, heap_factor 0.7, no refine, unsorted), drops itself and keeps nknn results.
        The N self-searches run as GPU batches.  A row with fewer than nknn results is padded with PAD_ID (the
        reference concatenates the rows, which only works when every row is full)."""
        if nknn <= 0:
            raise ValueError("nknn must be > 0")
        n = self._host.len
        out = np.full((n, nknn), N.PAD_ID, dtype=np.uint64)
        for lo in range(0, n, batch_docs):
            hi = min(n, lo + batch_docs)
            off, comps, vals = self._host.forward_csr(lo, hi)
            ids, _, counts = self.gpu.batch_search(off, comps, vals, nknn + 1, self._KNN_QUERY_CUT,
                                                   self._KNN_HEAP_FACTOR, 0, False)
            me = np.arange(lo, hi, dtype=np.uint64)[:, None]
            keep = (ids != me) & (np.arange(nknn + 1)[None, :] < counts[:, None])
            # stable compaction of every row to its first nknn kept entries
            order = np.argsort(~keep, axis=1, kind="stable")[:, :nknn]
            rows = np.take_along_axis(ids, order, axis=1)
            rows[~np.take_along_axis(keep, order, axis=1)] = N.PAD_ID
            out[lo:hi] = rows
        self._attach_knn(out)

    _KNN_SIDECAR = ".knn"  # <index file>.knn: the graph of an index saved with one (InvertedIndexBase.knn: Option<Knn>)

    def _write_knn_file(self, file_path: str) -> None:
        knn = self._host.knn
        tmp = "%s.tmp%d" % (file_path, os.getpid())
        with open(tmp, "wb") as f:
            f.write(self._KNN_MAGIC)
            f.write(np.array(knn.shape, dtype=np.uint64).tobytes())
            f.write(np.ascontiguousarray(knn).tobytes())
        os.replace(tmp, file_path)

    @classmethod
    def _read_knn_file(cls, file_path: str) -> np.ndarray:
        with open(file_path, "rb") as f:
            if f.read(8) != cls._KNN_MAGIC:
                raise OSError("%s is not a kNN file of this library" % file_path)
            n_vecs, dim = (int(x) for x in np.frombuffer(f.read(16), dtype=np.uint64))
            raw = f.read(n_vecs * dim * 8)
            if len(raw) != n_vecs * dim * 8:
                raise OSError("%s is truncated" % file_path)
            return np.frombuffer(raw, dtype=np.uint64).reshape(n_vecs, dim)

    def _save_host(self, index_file: str) -> None:
        """The index container + (like the reference, which serialises `knn: Option<Knn>` inside the index,
        src/inverted_index.rs:38-52) the attached kNN graph, as <index_file>.knn."""
        self._host.save(index_file)
        side = index_file + self._KNN_SIDECAR
        if self._host.knn is not None:
            self._write_knn_file(side)
        elif os.path.exists(side):
            os.remove(side)  # an index saved without a graph must not pick up a stale one

    @classmethod
    def _load_host(cls, index_file: str) -> HostIndex:
        host = HostIndex.load(index_file)
        side = index_file + cls._KNN_SIDECAR
        if os.path.exists(side):
            nb = cls._read_knn_file(side)
            if nb.shape[0] != host.len:
                raise OSError("%s does not belong to %s" % (side, index_file))
            host.set_knn(np.ascontiguousarray(nb))
        return host

    def save_knn(self, path: str) -> None:
        """Writes <path>.knn.seismic (Knn::serialize, src/inverted_index.rs:542-548).  Own flat format (the reference's
        byte encoding lives in vectorium): magic, n_vecs u64, dim u64, then n_vecs * dim u64 ids."""
        if self._host.knn is None:
            raise ValueError("No KNN graph is attached to the index.")  # PyValueError, src/pylib/mod.rs:260-264
        try:
            self._write_knn_file(path + ".knn.seismic")
        except OSError:
            raise
        except Exception as e:
            raise OSError(str(e))

    def load_knn(self, knn_path: str, nknn: Optional[int] = None) -> None:
        """Knn::new_from_serialized (src/inverted_index.rs:502-540): optionally keep only the first nknn neighbours."""
        nb = self._read_knn_file(knn_path)
        n_vecs, dim = nb.shape
        if n_vecs != self._host.len:
            raise ValueError("kNN file holds %d vectors, the index %d" % (n_vecs, self._host.len))
        if nknn is not None:
            if nknn > dim:  # assert in the reference, src/inverted_index.rs:513-516
                raise ValueError("The number of neighbors to include for each vector of the dataset can't be greater "
                                 "than the number of neighbours in the precomputed knn file.")
            nb = nb[:, :nknn]
        self._attach_knn(np.ascontiguousarray(nb))

    # -- GPU
    def to_device(self, device: int = 0) -> "GpuIndex":
        if self._gpu is None or self._device != device:
            self._device = device
            self._gpu = GpuIndex(self._host, device)
        return self._gpu

    @property
    def gpu(self) -> GpuIndex:
        return self.to_device(self._device)

    def _search_csr(self, offsets, comps, values, k, query_cut, heap_factor, n_knn, sorted):
        return self.gpu.batch_search(offsets, comps, values, k, query_cut, heap_factor, n_knn, sorted)

    @classmethod
    def _config(cls, n_postings, centroid_fraction, min_cluster_size, summary_energy, max_fraction, doc_cut,
                num_threads=0) -> N.BuildConfig:
        return make_config(n_postings=n_postings, centroid_fraction=centroid_fraction, min_cluster_size=min_cluster_size,
                           summary_energy=summary_energy, max_fraction=max_fraction, doc_cut=doc_cut,
                           comp_bits=cls._COMP_BITS, value_kind=cls._VALUE_KIND, n_threads=num_threads)

    def _apply_knn_args(self, nknn, knn_path):
        """KnnConfiguration of build (src/pylib/mod.rs:349-352, src/inverted_index.rs:654-685): a precomputed file
        wins over building; nknn then limits the neighbours read from it."""
        if knn_path:
            self.load_knn(knn_path, nknn if nknn else None)
        elif nknn:
            self.build_knn(nknn)
        return self


# ------------------------------------------------------------------------------------------ SeismicDataset
class SeismicDataset:
    """Growable dataset of (doc_id, {token: value}) with exact (brute-force) search."""
    _COMP_BITS = 16

    def __init__(self):
        self._token_to_id: Dict[str, int] = {}
        self._doc_ids: List[str] = []
        self._comps: List[np.ndarray] = []
        self._vals: List[np.ndarray] = []
        self._contents: List[Optional[str]] = []
        self._exact: Optional[Tuple[int, "_IndexBase"]] = None

    @property
    def len(self) -> int:
        return len(self._doc_ids)

    def __len__(self) -> int:
        return len(self._doc_ids)

    def add_document(self, doc_id: str, tokens, values, content: Optional[str] = None) -> None:
        tokens = [str(t) for t in np.asarray(tokens).tolist()]
        values = np.asarray(values, dtype=np.float32)
        if len(tokens) != len(values):
            raise ValueError("tokens and values must have the same length")
        for t in tokens:
            if t not in self._token_to_id:
                if len(self._token_to_id) + 1 >= 2 ** self._COMP_BITS:
                    raise ValueError("The number of different tokens exceeds 2^%d." % self._COMP_BITS)
                self._token_to_id[t] = len(self._token_to_id)
        c = np.array([self._token_to_id[t] for t in tokens], dtype=np.uint32)
        order = np.argsort(c, kind="stable")
        self._doc_ids.append(str(doc_id))
        self._comps.append(c[order])
        self._vals.append(values[order])
        self._contents.append(content)
        self._exact = None

    def get_doc_text(self, doc_id: str) -> Optional[str]:
        try:
            return self._contents[self._doc_ids.index(doc_id)]
        except ValueError:
            return None

    def _native(self) -> Dataset:
        return Dataset.from_lists(self._comps, self._vals, dim=max(1, len(self._token_to_id)))

    def _exact_index(self) -> "_IndexBase":
        """Exact search runs on the GPU over the f16 forward index (FlatIndex stand-in)."""
        if self._exact is None or self._exact[0] != len(self._doc_ids):
            host = HostIndex.build(self._native(), comp_bits=self._COMP_BITS)
            self._exact = (len(self._doc_ids), _IndexBase(host))
        return self._exact[1]

    def _resolve(self, tokens, values):
        pairs = sorted((self._token_to_id[str(t)], float(v)) for t, v in zip(np.asarray(tokens).tolist(), np.asarray(values).tolist())
                       if str(t) in self._token_to_id)
        return (np.array([p[0] for p in pairs], np.uint32), np.array([p[1] for p in pairs], np.float32))

    def search(self, query_id: str, query_components, query_values, k: int):
        return self.batch_search([query_id], [query_components], [query_values], k)[0]

    def batch_search(self, queries_ids, query_components, query_values, k: int, num_threads: int = 0):
        qs = [self._resolve(c, v) for c, v in zip(query_components, query_values)]
        off, qc, qv = csr_from_lists([q[0] for q in qs], [q[1] for q in qs])
        ids, scores, counts = self._exact_index().gpu.exact_search(off, qc, qv, k)
        out = []
        for qi, qid in enumerate(np.asarray(queries_ids).tolist()):
            n = int(counts[qi])
            out.append([(str(qid), float(scores[qi, r]), self._doc_ids[int(ids[qi, r])]) for r in range(n)])
        return out


class SeismicDatasetLV(SeismicDataset):
    _COMP_BITS = 32


# ------------------------------------------------------------------------------------------ SeismicIndex
class SeismicIndex(_IndexBase):
    """String-token / string-doc-id index (reference impl_seismic_index!, src/pylib/mod.rs:46-661)."""
    _COMP_BITS = 16
    _META_SUFFIX = ".meta.json"

    def __init__(self, host: HostIndex, doc_ids: Optional[List[str]], token_to_id: Dict[str, int],
                 contents: Optional[List[Optional[str]]] = None):
        super().__init__(host)
        self._doc_ids = doc_ids
        self._token_to_id = token_to_id
        self._contents = contents
        self._doc_pos: Optional[Dict[str, int]] = None

    # -- construction
    @classmethod
    def build(cls, input_path: str, n_postings: int = 3500, centroid_fraction: float = 0.1, min_cluster_size: int = 2,
              summary_energy: float = 0.4, max_fraction: float = 1.5, doc_cut: int = 15, nknn: int = 0,
              knn_path: Optional[str] = None, batched_indexing: Optional[int] = None,
              input_token_to_id_map: Optional[Dict[str, int]] = None, load_content: bool = True, num_threads: int = 0):
        try:
            token_to_id, doc_ids, contents, comps, vals = _read_collection(
                input_path, dict(input_token_to_id_map) if input_token_to_id_map else None, load_content,
                2 ** cls._COMP_BITS)
        except FileNotFoundError as e:
            raise OSError(str(e))
        ds = Dataset.from_lists(comps, vals, dim=max(1, len(token_to_id)))
        cfg = cls._config(n_postings, centroid_fraction, min_cluster_size, summary_energy, max_fraction, doc_cut, num_threads)
        return cls(cls._build_host(ds, cfg), doc_ids, token_to_id, contents)._apply_knn_args(nknn, knn_path)

    @classmethod
    def build_from_dataset(cls, dataset: SeismicDataset, n_postings: int = 3500, centroid_fraction: float = 0.1,
                           min_cluster_size: int = 2, summary_energy: float = 0.4, max_fraction: float = 1.5,
                           doc_cut: int = 15, nknn: int = 0, knn_path: Optional[str] = None,
                           batched_indexing: Optional[int] = None, num_threads: int = 0):
        cfg = cls._config(n_postings, centroid_fraction, min_cluster_size, summary_energy, max_fraction, doc_cut, num_threads)
        return cls(cls._build_host(dataset._native(), cfg), list(dataset._doc_ids), dict(dataset._token_to_id),
                   list(dataset._contents))._apply_knn_args(nknn, knn_path)

    @classmethod
    def _build_host(cls, ds: Dataset, cfg: N.BuildConfig) -> HostIndex:
        return HostIndex.build(ds, cfg)

    # -- persistence: <path>.index.seismic (flat container, see csrc/host/io.cpp) + <path>.index.seismic.meta.json
    def save(self, path: str) -> None:
        try:
            self._save_host(path + ".index.seismic")
            with open(path + ".index.seismic" + self._META_SUFFIX, "w", encoding="utf-8") as f:
                json.dump({"doc_ids": self._doc_ids, "token_to_id": self._token_to_id, "contents": self._contents}, f)
        except OSError:
            raise
        except Exception as e:  # PyIOError in the reference
            raise OSError(str(e))

    @classmethod
    def load(cls, index_path: str):
        try:
            host = cls._load_host(index_path)
            meta = {}
            if os.path.exists(index_path + cls._META_SUFFIX):
                with open(index_path + cls._META_SUFFIX, "r", encoding="utf-8") as f:
                    meta = json.load(f)
        except OSError:
            raise
        except Exception as e:
            raise OSError(str(e))
        return cls(host, meta.get("doc_ids"), meta.get("token_to_id", {}), meta.get("contents"))

    # -- search
    def _resolve(self, tokens, values) -> Tuple[np.ndarray, np.ndarray]:
        """resolve_query_tokens (src/inverted_index_wrapper.rs:75-91): unknown tokens dropped, sorted by component."""
        t2i = self._token_to_id
        pairs = sorted((t2i[str(t)], float(v)) for t, v in zip(np.asarray(tokens).tolist(), np.asarray(values).tolist())
                       if str(t) in t2i)
        return (np.array([p[0] for p in pairs], np.uint32), np.array([p[1] for p in pairs], np.float32))

    def _doc_name(self, idx: int) -> str:
        return self._doc_ids[idx] if self._doc_ids is not None else str(idx)

    def search(self, query_id: str, query_components, query_values, k: int, query_cut: int, heap_factor: float,
               n_knn: int = 0, sorted: bool = True) -> List[Tuple[str, float, str]]:
        return self.batch_search([query_id], [query_components], [query_values], k, query_cut, heap_factor, n_knn, sorted)[0]

    def batch_search(self, queries_ids, query_components: Sequence, query_values: Sequence, k: int, query_cut: int,
                     heap_factor: float, n_knn: int = 0, sorted: bool = True, num_threads: int = 0):
        """list[list[(query_id, score, doc_id)]], one list per query, in input order (the reference's par_bridge
        leaves the order unspecified, src/pylib/mod.rs:629-652)."""
        qs = [self._resolve(c, v) for c, v in zip(query_components, query_values)]
        off, qc, qv = csr_from_lists([q[0] for q in qs], [q[1] for q in qs])
        ids, scores, counts = self._search_csr(off, qc, qv, k, query_cut, heap_factor, n_knn, sorted)
        out = []
        for qi, qid in enumerate(np.asarray(queries_ids).tolist()):
            n = int(counts[qi])
            out.append([(str(qid), float(scores[qi, r]), self._doc_name(int(ids[qi, r]))) for r in range(n)])
        return out

    def get_doc_text(self, doc_id: str) -> Optional[str]:
        if self._doc_ids is None or self._contents is None:
            return None
        if self._doc_pos is None:
            self._doc_pos = {d: i for i, d in enumerate(self._doc_ids)}
        i = self._doc_pos.get(doc_id)
        return None if i is None else self._contents[i]


class SeismicIndexLV(SeismicIndex):
    """Large vocabulary: u32 components (reference impl_seismic_index!(…, u32, …), src/pylib/mod.rs:1160-1166)."""
    _COMP_BITS = 32


class SeismicIndexDotVByte(SeismicIndex):
    """Built as a standard u16/f16 index, then the forward index is converted to the DotVByte encoding
    (reference src/pylib/dotvbyte.rs:195-213, src/inverted_index.rs:237-275)."""

    @classmethod
    def _build_host(cls, ds: Dataset, cfg: N.BuildConfig) -> HostIndex:
        host = HostIndex.build(ds, cfg)
        return host.convert_to_dotvbyte()

    # the reference class has no kNN methods (src/pylib/dotvbyte.rs:101-112)
    def build_knn(self, nknn: int, batch_docs: int = 0) -> None:
        raise AttributeError("SeismicIndexDotVByte has no build_knn")

    def load_knn(self, knn_path: str, nknn: Optional[int] = None) -> None:
        raise AttributeError("SeismicIndexDotVByte has no load_knn")

    def _apply_knn_args(self, nknn, knn_path):
        if nknn or knn_path:
            raise ValueError("SeismicIndexDotVByte does not support kNN graphs")
        return self


# ------------------------------------------------------------------------------------------ Raw indexes
class SeismicIndexRaw(_IndexBase):
    """Integer components, integer doc ids, datasets in the seismic inner .bin format (src/pylib/mod.rs:663-1151)."""
    _COMP_BITS = 16

    @classmethod
    def build(cls, input_file: str, n_postings: int = 3500, centroid_fraction: float = 0.1, min_cluster_size: int = 2,
              summary_energy: float = 0.4, max_fraction: float = 1.5, doc_cut: int = 15, nknn: int = 0,
              knn_path: Optional[str] = None, batched_indexing: Optional[int] = None):
        ds = Dataset.read_bin(input_file)
        cfg = cls._config(n_postings, centroid_fraction, min_cluster_size, summary_energy, max_fraction, doc_cut)
        return cls(HostIndex.build(ds, cfg))._apply_knn_args(nknn, knn_path)

    def save(self, path: str) -> None:
        self._save_host(path + ".index.seismic")

    @classmethod
    def load(cls, index_path: str):
        return cls(cls._load_host(index_path))

    def search(self, query_components, query_values, k: int, query_cut: int, heap_factor: float, n_knn: int,
               sorted: bool) -> List[Tuple[float, int]]:
        qc = np.asarray(query_components).astype(np.uint32)
        qv = np.asarray(query_values, dtype=np.float32)
        ids, scores, counts = self._search_csr(np.array([0, len(qc)], np.uint64), qc, qv, k, query_cut, heap_factor,
                                               n_knn, sorted)
        return [(float(scores[0, r]), int(ids[0, r])) for r in range(int(counts[0]))]

    def batch_search(self, query_path: str, k: int, query_cut: int, heap_factor: float, n_knn: int, sorted: bool,
                     num_threads: int = 0) -> List[List[Tuple[float, int]]]:
        q = Dataset.read_bin(query_path)
        ids, scores, counts = self._search_csr(q.offsets, q.comps, q.values, k, query_cut, heap_factor, n_knn, sorted)
        return [[(float(scores[qi, r]), int(ids[qi, r])) for r in range(int(counts[qi]))] for qi in range(len(q))]

    def write_results_tsv(self, results: List[List[Tuple[float, int]]], path: str) -> None:
        """`query_id\\tdoc\\trank\\tscore` (reference src/bin/perf_inverted_index.rs:223-235)."""
        with open(path, "w") as f:
            for qi, res in enumerate(results):
                for rank, (score, doc) in enumerate(res):
                    f.write("%d\t%d\t%d\t%s\n" % (qi, doc, rank + 1, repr(float(np.float32(score)))))


class SeismicIndexRawLV(SeismicIndexRaw):
    _COMP_BITS = 32
