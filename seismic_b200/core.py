"""Low-level Python objects over the C ABI: Dataset, HostIndex (CPU build / persistence) and GpuIndex
(the HBM image + batched search).  The `seismic`-compatible classes live in seismic_b200/api.py."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np

from . import _native as N


def csr_from_lists(components: Sequence[np.ndarray], values: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Concatenate per-vector arrays into (offsets u64, comps u32, values f32)."""
    n = len(components)
    offsets = np.zeros(n + 1, dtype=np.uint64)
    if n:
        offsets[1:] = np.cumsum([len(c) for c in components], dtype=np.uint64)
    comps = np.concatenate([np.asarray(c, dtype=np.uint32) for c in components]) if n and offsets[-1] else np.empty(0, np.uint32)
    vals = np.concatenate([np.asarray(v, dtype=np.float32) for v in values]) if n and offsets[-1] else np.empty(0, np.float32)
    return offsets, np.ascontiguousarray(comps, dtype=np.uint32), np.ascontiguousarray(vals, dtype=np.float32)


class Dataset:
    """Sparse dataset (CSR, u32 components, f32 values) owned by the native library."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            N.hlib().shost_dataset_destroy(self._h)
            self._h = C.c_void_p(0)

    @staticmethod
    def from_csr(offsets, comps, values, dim: int) -> "Dataset":
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        comps = np.ascontiguousarray(comps, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.float32)
        if len(comps) != len(values) or int(offsets[-1]) != len(comps):
            raise ValueError("inconsistent CSR arrays")
        out = C.c_void_p()
        N.hcheck(N.hlib().shost_dataset_create(len(offsets) - 1, dim, N.ptr(offsets), N.ptr(comps), N.ptr(values), C.byref(out)))
        return Dataset(out.value)

    @staticmethod
    def from_lists(components, values, dim: Optional[int] = None) -> "Dataset":
        off, c, v = csr_from_lists(components, values)
        if dim is None:
            dim = int(c.max()) + 1 if len(c) else 0
        return Dataset.from_csr(off, c, v, dim)

    @staticmethod
    def read_bin(path: str) -> "Dataset":
        out = C.c_void_p()
        N.hcheck(N.hlib().shost_dataset_read_bin(str(path).encode(), C.byref(out)))
        return Dataset(out.value)

    def write_bin(self, path: str) -> None:
        N.hcheck(N.hlib().shost_dataset_write_bin(self._h, str(path).encode()))

    @staticmethod
    def synth_config(n_docs: int, dim: int = 30522, seed: int = 20260517, n_topics: Optional[int] = None, **kw) -> N.SynthConfig:
        cfg = N.SynthConfig()
        N.hlib().shost_default_synth(C.byref(cfg))
        cfg.n_docs, cfg.dim, cfg.seed = n_docs, dim, seed
        # keep ~2150 documents per topic at every scale (4096 topics at 8.8 M documents)
        cfg.n_topics = n_topics if n_topics else int(min(4096, max(16, n_docs // 2150)))
        for k, v in kw.items():
            setattr(cfg, k, v)
        return cfg

    @staticmethod
    def synth_documents(cfg: N.SynthConfig) -> "Dataset":
        out = C.c_void_p()
        N.hcheck(N.hlib().shost_synth_documents(C.byref(cfg), C.byref(out)))
        return Dataset(out.value)

    @staticmethod
    def synth_queries(cfg: N.SynthConfig, n_queries: int) -> "Dataset":
        out = C.c_void_p()
        N.hcheck(N.hlib().shost_synth_queries(C.byref(cfg), n_queries, C.byref(out)))
        return Dataset(out.value)

    def __len__(self) -> int:
        return int(N.hlib().shost_dataset_len(self._h))

    @property
    def dim(self) -> int:
        return int(N.hlib().shost_dataset_dim(self._h))

    @property
    def nnz(self) -> int:
        return int(N.hlib().shost_dataset_nnz(self._h))

    # borrowed views (valid while `self` is alive)
    @property
    def offsets(self) -> np.ndarray:
        return N.np_view(N.hlib().shost_dataset_offsets(self._h), len(self) + 1, np.uint64)

    @property
    def comps(self) -> np.ndarray:
        return N.np_view(N.hlib().shost_dataset_comps(self._h), self.nnz, np.uint32)

    @property
    def values(self) -> np.ndarray:
        return N.np_view(N.hlib().shost_dataset_values(self._h), self.nnz, np.float32)

    def vector(self, i: int) -> Tuple[np.ndarray, np.ndarray]:
        o = self.offsets
        return self.comps[int(o[i]):int(o[i + 1])], self.values[int(o[i]):int(o[i + 1])]


def make_config(n_postings=3500, centroid_fraction=0.1, min_cluster_size=2, summary_energy=0.4, max_fraction=1.5,
                doc_cut=15, comp_bits=16, value_kind=N.VAL_F16, n_threads=0, **extra) -> N.BuildConfig:
    """ShostBuildConfig with the Python defaults of the reference (src/pylib/mod.rs:329)."""
    cfg = N.BuildConfig()
    N.hlib().shost_default_config(C.byref(cfg))
    cfg.n_postings, cfg.centroid_fraction, cfg.min_cluster_size = n_postings, centroid_fraction, min_cluster_size
    cfg.summary_energy, cfg.max_fraction, cfg.doc_cut = summary_energy, max_fraction, doc_cut
    cfg.comp_bits, cfg.value_kind, cfg.n_threads = comp_bits, value_kind, n_threads
    for k, v in extra.items():
        setattr(cfg, k, v)
    return cfg


class HostIndex:
    """Logical Seismic index in host memory (built on the CPU or mmap-loaded)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)
        self._view = N.IndexView()
        N.hcheck(N.hlib().shost_index_view(self._h, C.byref(self._view)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            N.hlib().shost_index_destroy(self._h)
            self._h = C.c_void_p(0)

    @staticmethod
    def build(dataset: Dataset, config: Optional[N.BuildConfig] = None, **params) -> "HostIndex":
        cfg = config if config is not None else make_config(**params)
        out = C.c_void_p()
        N.hcheck(N.hlib().shost_index_build(dataset._h, C.byref(cfg), C.byref(out)))
        return HostIndex(out.value)

    @staticmethod
    def load(path: str) -> "HostIndex":
        out = C.c_void_p()
        N.hcheck(N.hlib().shost_index_load(str(path).encode(), C.byref(out)))
        return HostIndex(out.value)

    def save(self, path: str) -> None:
        N.hcheck(N.hlib().shost_index_save(self._h, str(path).encode()))

    def convert_to_dotvbyte(self) -> "HostIndex":
        """Same posting lists over a DotVByte forward index (reference src/pylib/dotvbyte.rs:195-213)."""
        out = C.c_void_p()
        N.hcheck(N.hlib().shost_index_convert_dotvbyte(self._h, C.byref(out)))
        return HostIndex(out.value)

    # -- kNN graph (reference Knn{n_vecs, dim, neighbours}, src/inverted_index.rs:430-434)
    def set_knn(self, neighbours: Optional[np.ndarray]) -> None:
        """Attach (or, with None, drop) a kNN graph: uint64 [len, dim] document ids, PAD_ID = no neighbour."""
        if neighbours is None:
            self._knn = None
            self._view.knn_dim, self._view.knn_neighbours = 0, None
            return
        nb = np.ascontiguousarray(neighbours, dtype=np.uint64)
        if nb.ndim != 2 or nb.shape[0] != self.len or nb.shape[1] == 0:
            raise ValueError("kNN graph must have shape [len, dim > 0]")
        self._knn = nb  # keeps the buffer the view points to alive
        self._view.knn_dim, self._view.knn_neighbours = nb.shape[1], N.ptr(nb)

    @property
    def knn(self) -> Optional[np.ndarray]:
        return getattr(self, "_knn", None)

    @property
    def value_kind(self) -> int:
        return int(self._view.value_kind)

    @property
    def view(self) -> N.IndexView:
        return self._view

    @property
    def len(self) -> int:
        return int(self._view.n_docs)

    @property
    def dim(self) -> int:
        return int(self._view.dim)

    @property
    def nnz(self) -> int:
        return int(N.hlib().shost_index_nnz(self._h))

    @property
    def comp_bits(self) -> int:
        return int(self._view.comp_bits)

    def space_usage(self) -> dict:
        b = (C.c_uint64 * 6)()
        N.hcheck(N.hlib().shost_index_space_usage(self._h, b))
        return dict(zip(["forward", "packed_postings", "block_offsets", "summaries", "knn", "total"], map(int, b)))

    def get_doc(self, i: int) -> Tuple[np.ndarray, np.ndarray]:
        cap = 65536
        comps = np.empty(cap, np.uint32)
        vals = np.empty(cap, np.float32)
        n = C.c_uint32()
        N.hcheck(N.hlib().shost_index_get_doc(self._h, i, N.ptr(comps), N.ptr(vals), cap, C.byref(n)))
        return comps[: n.value].copy(), vals[: n.value].copy()

    def forward_csr(self, lo: int = 0, hi: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """Documents [lo, hi) of the forward index as CSR queries: (offsets u64, components u32, values f32) — the
        vectors `Knn::new` searches the index with (src/inverted_index.rs:463-470).  Plain encodings only."""
        v = self._view
        hi = self.len if hi is None else hi
        if v.value_kind == N.VAL_DOTVBYTE:
            raise NotImplementedError("forward_csr: DotVByte forward index")
        fo = N.np_view(v.fwd_offsets, self.len + 1, np.uint64)
        a, b = int(fo[lo]), int(fo[hi])
        comps = N.np_view(v.fwd_comps, int(fo[-1]), np.uint16 if v.comp_bits == 16 else np.uint32)[a:b].astype(np.uint32)
        kind, total = int(v.value_kind), int(fo[-1])
        if kind == N.VAL_F16:
            vals = N.np_view(v.fwd_values, total, np.float16)[a:b].astype(np.float32)
        elif kind == N.VAL_BF16:
            vals = (N.np_view(v.fwd_values, total, np.uint16)[a:b].astype(np.uint32) << 16).view(np.float32)
        elif kind == N.VAL_F32:
            vals = N.np_view(v.fwd_values, total, np.float32)[a:b].copy()
        else:
            code = N.np_view(v.fwd_values, total, np.uint8 if kind == N.VAL_FIXEDU8 else np.uint16)[a:b]
            vals = code.astype(np.float32) * np.float32(v.value_scale)
        return (fo[lo:hi + 1] - fo[lo]).astype(np.uint64), comps, vals

    # numpy views of the logical arrays (tests, tools)
    def arrays(self) -> dict:
        v = self._view
        dim, n = int(v.dim), int(v.n_docs)
        lps = N.np_view(v.list_post_start, dim + 1, np.uint64)
        lbs = N.np_view(v.list_blk_start, dim + 1, np.uint64)
        lss = N.np_view(v.list_sc_start, dim + 1, np.uint64)
        les = N.np_view(v.list_ent_start, dim + 1, np.uint64)
        fo = N.np_view(v.fwd_offsets, n + 1, np.uint64)
        return {
            "fwd_offsets": fo,
            "fwd_comps": N.np_view(v.fwd_comps, int(fo[-1]) if v.fwd_comps else 0,
                                   np.uint16 if v.comp_bits == 16 else np.uint32),
            "list_post_start": lps, "postings": N.np_view(v.postings, int(lps[-1]), np.uint64),
            "list_blk_start": lbs, "blk_post_off": N.np_view(v.blk_post_off, int(lbs[-1]) + dim, np.uint32),
            "blk_min": N.np_view(v.blk_min, int(lbs[-1]), np.float32),
            "blk_quant": N.np_view(v.blk_quant, int(lbs[-1]), np.float32),
            "list_sc_start": lss, "sc_comp": N.np_view(v.sc_comp, int(lss[-1]), np.uint32),
            "list_ent_start": les, "sc_run_off": N.np_view(v.sc_run_off, int(lss[-1]) + dim, np.uint32),
            "ent_blk": N.np_view(v.ent_blk, int(les[-1]), np.uint16),
            "ent_code": N.np_view(v.ent_code, int(les[-1]), np.uint8),
        }


class GpuIndex:
    """HBM image of a HostIndex on one CUDA device + the batched search entry points."""

    def __init__(self, host: HostIndex, device: int = 0):
        self._h = C.c_void_p()
        N.check(N.lib().sgpu_index_create(C.byref(host.view), device, C.byref(self._h)))
        self.device = device
        self.dim = host.dim
        self.len = host.len
        self.last_stats: dict = {}
        # tuning knobs for experiments: SEISMIC_B200_OPTS="hq=2,hq_wave_docs=512" (never changes results)
        import os
        for kv in filter(None, os.environ.get("SEISMIC_B200_OPTS", "").split(",")):
            name, _, val = kv.partition("=")
            self.set_option(name.strip(), int(val))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().sgpu_index_destroy(self._h)
            self._h = C.c_void_p(0)

    @property
    def device_bytes(self) -> int:
        return int(N.lib().sgpu_index_device_bytes(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        """Run on the caller's CUDA stream (e.g. torch.cuda.current_stream().cuda_stream); 0 = private stream."""
        N.check(N.lib().sgpu_index_set_stream(self._h, C.c_void_p(cuda_stream)))

    def set_option(self, name: str, value: int) -> None:
        N.check(N.lib().sgpu_index_set_option(self._h, name.encode(), int(value)))

    def set_knn(self, neighbours: Optional[np.ndarray]) -> None:
        """(Re)upload the kNN graph used by n_knn > 0 searches: uint64 [len, dim] document ids; None drops it."""
        if neighbours is None:
            N.check(N.lib().sgpu_index_set_knn(self._h, None, 0))
            return
        nb = np.ascontiguousarray(neighbours, dtype=np.uint64)
        if nb.ndim != 2 or nb.shape[0] != self.len:
            raise ValueError("kNN graph must have shape [len, dim]")
        N.check(N.lib().sgpu_index_set_knn(self._h, N.ptr(nb), nb.shape[1]))

    @staticmethod
    def _params(k, query_cut, heap_factor, n_knn, first_sorted) -> N.SearchParams:
        if k <= 0:
            raise ValueError("k must be > 0")
        return N.SearchParams(int(k), int(query_cut), float(heap_factor), int(n_knn), 1 if first_sorted else 0)

    def batch_search(self, offsets, comps, values, k, query_cut, heap_factor, n_knn=0, first_sorted=True, out=None):
        """Host buffers in, host buffers out (H2D/D2H inside the call). Returns ids[nq,k], scores[nq,k], counts[nq].
        `out` = (ids, scores, counts) arrays to fill instead of fresh ones; page-locked arrays (`pinned_array`), for
        inputs and outputs alike, are read / written by DMA without the staging copy."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        comps = np.ascontiguousarray(comps, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.float32)
        nq = len(offsets) - 1
        if out is not None:
            ids, scores, counts = out
            if (ids.shape != (nq, k) or ids.dtype != np.uint64 or scores.shape != (nq, k) or scores.dtype != np.float32
                    or counts.shape != (nq,) or counts.dtype != np.uint32
                    or not (ids.flags.c_contiguous and scores.flags.c_contiguous and counts.flags.c_contiguous)):
                raise ValueError("out = (uint64[nq,k], float32[nq,k], uint32[nq]) C-contiguous arrays")
        else:
            ids = np.empty((nq, k), dtype=np.uint64)
            scores = np.empty((nq, k), dtype=np.float32)
            counts = np.empty(nq, dtype=np.uint32)
        qb = N.QueryBatch(nq, N.ptr(offsets), N.ptr(comps), N.ptr(values))
        p = self._params(k, query_cut, heap_factor, n_knn, first_sorted)
        st = N.SearchStats()
        N.check(N.lib().sgpu_batch_search(self._h, C.byref(qb), C.byref(p), N.ptr(ids), N.ptr(scores), N.ptr(counts), C.byref(st)))
        self.last_stats = st.as_dict()
        return ids, scores, counts

    def batch_search_device(self, d_offsets, d_comps, d_values, nq, k, query_cut, heap_factor, d_ids, d_scores,
                            d_counts, n_knn=0, first_sorted=True) -> dict:
        """All arguments are raw device pointers (ints, e.g. torch.Tensor.data_ptr()) on this index's device."""
        qb = N.QueryBatch(nq, d_offsets, d_comps, d_values)
        p = self._params(k, query_cut, heap_factor, n_knn, first_sorted)
        st = N.SearchStats()
        N.check(N.lib().sgpu_batch_search_device(self._h, C.byref(qb), C.byref(p), d_ids, d_scores, d_counts, C.byref(st)))
        self.last_stats = st.as_dict()
        return self.last_stats

    def exact_search(self, offsets, comps, values, k):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        comps = np.ascontiguousarray(comps, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.float32)
        nq = len(offsets) - 1
        ids = np.empty((nq, k), dtype=np.uint64)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.uint32)
        qb = N.QueryBatch(nq, N.ptr(offsets), N.ptr(comps), N.ptr(values))
        ms = C.c_float()
        N.check(N.lib().sgpu_exact_search(self._h, C.byref(qb), int(k), N.ptr(ids), N.ptr(scores), N.ptr(counts), C.byref(ms)))
        self.last_stats = {"ms_exact": ms.value}
        return ids, scores, counts


class GpuGroup:
    """The index replicated on several GPUs of one box; one batch is split across them and the result tuples are
    gathered on the first device with one NCCL group (sgpu_group_*, single process — what a Rust host would call)."""

    def __init__(self, host: HostIndex, devices: Sequence[int]):
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        self._h = C.c_void_p()
        N.check(N.lib().sgpu_group_create(C.byref(host.view), devs, len(devices), C.byref(self._h)))
        self.devices = list(devices)
        self.last_stats: dict = {}

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            N.lib().sgpu_group_destroy(self._h)
            self._h = C.c_void_p(0)

    def __len__(self) -> int:
        return int(N.lib().sgpu_group_size(self._h))

    def set_knn(self, neighbours: Optional[np.ndarray]) -> None:
        if neighbours is None:
            N.check(N.lib().sgpu_group_set_knn(self._h, None, 0))
            return
        nb = np.ascontiguousarray(neighbours, dtype=np.uint64)
        N.check(N.lib().sgpu_group_set_knn(self._h, N.ptr(nb), nb.shape[1]))

    def batch_search(self, offsets, comps, values, k, query_cut, heap_factor, n_knn=0, first_sorted=True):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        comps = np.ascontiguousarray(comps, dtype=np.uint32)
        values = np.ascontiguousarray(values, dtype=np.float32)
        nq = len(offsets) - 1
        ids = np.empty((nq, k), dtype=np.uint64)
        scores = np.empty((nq, k), dtype=np.float32)
        counts = np.empty(nq, dtype=np.uint32)
        qb = N.QueryBatch(nq, N.ptr(offsets), N.ptr(comps), N.ptr(values))
        p = GpuIndex._params(k, query_cut, heap_factor, n_knn, first_sorted)
        st = N.SearchStats()
        ms = C.c_float()
        N.check(N.lib().sgpu_group_batch_search(self._h, C.byref(qb), C.byref(p), N.ptr(ids), N.ptr(scores),
                                                N.ptr(counts), C.byref(st), C.byref(ms)))
        self.last_stats = st.as_dict()
        self.last_stats["ms_gather"] = ms.value
        return ids, scores, counts


class _PinnedBlock:
    def __init__(self, nbytes: int):
        self.ptr = C.c_void_p()
        N.check(N.lib().sgpu_host_alloc(int(nbytes), C.byref(self.ptr)))

    def __del__(self):
        if getattr(self, "ptr", None) and self.ptr.value:
            N.lib().sgpu_host_free(self.ptr)
            self.ptr = C.c_void_p(0)


def pinned_array(shape, dtype) -> np.ndarray:
    """A numpy array in page-locked host memory (sgpu_host_alloc): the search calls DMA straight from / into it."""
    dtype = np.dtype(dtype)
    shape = (int(shape),) if np.isscalar(shape) else tuple(int(x) for x in shape)
    nbytes = int(np.prod(shape, dtype=np.int64)) * dtype.itemsize
    block = _PinnedBlock(max(nbytes, 1))
    buf = (C.c_uint8 * max(nbytes, 1)).from_address(block.ptr.value)
    buf._block = block  # the memory lives as long as any array derived from `buf`
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape, dtype=np.int64))).reshape(shape)


def recall_at_k(exact_ids: np.ndarray, exact_counts: np.ndarray, run_ids: np.ndarray, run_counts: np.ndarray) -> float:
    """accuracy@k = sum_q |gt_q ∩ run_q| / sum_q |gt_q| (reference scripts/run_experiments.py:287-309)."""
    hit = 0
    tot = 0
    for q in range(len(exact_counts)):
        gt = set(exact_ids[q, : int(exact_counts[q])].tolist())
        rn = set(run_ids[q, : int(run_counts[q])].tolist())
        hit += len(gt & rn)
        tot += len(gt)
    return hit / tot if tot else 1.0
