"""The C++ oracle against an independent pure-Python restatement of the reference algorithm (tests/py_reference.py)
on small random indexes: identical ids, bit-identical f32 scores, identical number of evaluated blocks.  Two
restatements written separately from the Rust sources agreeing on every discrete decision is the strongest pin this
repository can put on the oracle beyond the reference's own known-answer tests (no Rust toolchain here)."""
import numpy as np
import pytest

from seismic_b200 import Dataset, HostIndex
from py_reference import PyIndex, PAD, decode_dotvbyte


def random_index(seed, n_docs=300, dim=60, **build):
    rng = np.random.default_rng(seed)
    comps, vals = [], []
    for _ in range(n_docs):
        nnz = int(rng.integers(0, 14))  # empty documents included
        c = np.sort(rng.choice(dim, size=nnz, replace=False)).astype(np.uint32)
        comps.append(c)
        vals.append(rng.choice([0.25, 0.5, 1.0, 1.5, 2.0, 3.0], size=nnz).astype(np.float32) if seed % 2 else
                    rng.random(nnz, dtype=np.float32) * 3)
    return HostIndex.build(Dataset.from_lists(comps, vals, dim=dim), **build), rng


@pytest.mark.parametrize("seed,build", [
    (1, dict(n_postings=40, centroid_fraction=0.2)),       # quantised values: exact score ties everywhere
    (2, dict(n_postings=40, centroid_fraction=0.2)),
    (3, dict(n_postings=8, centroid_fraction=0.5, max_fraction=2.0, summary_energy=0.3)),   # heavy pruning
    (4, dict(n_postings=100, centroid_fraction=0.05, min_cluster_size=1, summary_energy=0.9)),
])
@pytest.mark.parametrize("k,cut,hf,srt,n_knn", [
    (5, 2, 0.8, True, 0), (5, 2, 0.8, False, 0), (1, 3, 1.0, True, 0), (10, 6, 0.5, True, 0), (3, 1, 1.3, False, 0),
    (5, 2, 0.9, True, 3), (8, 3, 0.8, False, 5)])
def test_oracle_matches_python_restatement(oracle_mod, seed, build, k, cut, hf, srt, n_knn):
    index, rng = random_index(seed, **build)
    if n_knn:
        g = rng.integers(0, index.len, size=(index.len, 4), dtype=np.uint64)
        g[rng.random(g.shape) < 0.2] = PAD
        index.set_knn(g)
    py = PyIndex(index)
    queries = []
    for _ in range(25):
        nnz = int(rng.integers(1, 9))
        c = np.sort(rng.choice(index.dim, size=nnz, replace=False)).astype(np.uint32)
        v = (rng.choice([0.5, 1.0, 1.0, 2.0], size=nnz) if seed % 2 else rng.random(nnz) * 2).astype(np.float32)
        queries.append((c, v))
    off = np.cumsum([0] + [len(c) for c, _ in queries]).astype(np.uint64)
    qc = np.concatenate([c for c, _ in queries])
    qv = np.concatenate([v for _, v in queries])
    ids, scores, counts, st = oracle_mod.batch_search(index.view, off, qc, qv, k, cut, hf, n_knn=n_knn,
                                                      first_sorted=srt, n_threads=1)
    evaluated = 0
    for i, (c, v) in enumerate(queries):
        p_ids, p_scores, ev = py.search(c, v, k, cut, hf, n_knn=n_knn, first_sorted=srt)
        evaluated += ev
        assert counts[i] == len(p_ids), (i, counts[i], p_ids)
        assert ids[i, : counts[i]].tolist() == p_ids, (i, ids[i], p_ids)
        assert np.array_equal(scores[i, : counts[i]], np.array(p_scores, dtype=np.float32)), i
    assert st["blocks_evaluated"] == evaluated


def test_dotvbyte_format_and_search_against_python_decoder(oracle_mod):
    """DotVByte (row a11): an independent decoder of the byte stream recovers every component exactly and every value
    within half a quantisation step; the oracle's fused decode + search on the packed index equals the Python
    restatement run on the decoded vectors (ids, score bits)."""
    index, rng = random_index(7, n_docs=400, dim=3000, n_postings=30, centroid_fraction=0.2)  # gaps >= 256 occur
    vb = index.convert_to_dotvbyte()
    off, comps, vals, codes = decode_dotvbyte(vb)
    o0, c0, v0 = index.forward_csr()
    # SeismicIndexDotVByte.get(id): the host decoder of the library agrees with the independent one
    for d in (0, 7, index.len - 1):
        gc, gv = vb.get_doc(d)
        assert np.array_equal(gc, comps[int(off[d]):int(off[d + 1])]) and np.array_equal(gv, vals[int(off[d]):int(off[d + 1])])
    assert np.array_equal(off, o0) and np.array_equal(comps, c0)
    scale = float(vb.view.value_scale)
    assert np.all(np.abs(vals - v0) <= scale / 2 + 1e-6)
    assert (np.diff(c0.astype(np.int64))[np.diff(c0.astype(np.int64)) > 0] >= 256).any(), "no 2-byte gap exercised"

    class Decoded(PyIndex):  # the Python search over the decoded vectors, postings of the packed index
        def __init__(self, host, csr, codes, scale):
            self.a = host.arrays()
            self.n_docs, self.dim = host.len, host.dim
            self.off, self.comps, _ = csr
            # DotVByte scores are the fixed-point dot product scaled once per document, carried at 2^-24
            # (oracle_search.cpp, doc_score_vbyte): per-component "value" = code * 2^-24, final factor scale * 2^24
            self.vals = (codes.astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
            self.s24 = np.float32(np.float32(scale) * np.float32(2.0 ** 24))
            self.fo = self.a["fwd_offsets"].astype(np.int64)
            self.knn = None
            self._el = csr[0].astype(np.int64)

        def doc_of(self, start):  # posting start = byte offset / 16 of the packed stream
            return int(np.searchsorted(self.fo, start * 16, side="right")) - 1

        def doc_score(self, q, start, ln):
            return np.float32(super().doc_score(q, int(self._el[self.doc_of(start)]), ln) * self.s24)

    py = Decoded(vb, (off, comps, vals), codes, vb.view.value_scale)
    queries = []
    for _ in range(20):
        nnz = int(rng.integers(1, 9))
        c = np.sort(rng.choice(index.dim, size=nnz, replace=False)).astype(np.uint32)
        queries.append((c, (rng.random(nnz) * 2).astype(np.float32)))
    qoff = np.cumsum([0] + [len(c) for c, _ in queries]).astype(np.uint64)
    qc = np.concatenate([c for c, _ in queries])
    qv = np.concatenate([v for _, v in queries])
    ids, scores, counts, _ = oracle_mod.batch_search(vb.view, qoff, qc, qv, 5, 3, 0.8, first_sorted=True, n_threads=1)
    for i, (c, v) in enumerate(queries):
        p_ids, p_scores, _ = py.search(c, v, 5, 3, 0.8, first_sorted=True)
        assert counts[i] == len(p_scores) and ids[i, : counts[i]].tolist() == p_ids, i
        assert np.array_equal(scores[i, : counts[i]], np.array(p_scores, dtype=np.float32)), i


@pytest.mark.parametrize("seed", range(10, 22))
def test_oracle_matches_python_restatement_random_configs(oracle_mod, seed):
    """Random build parameters, k, query_cut, heap_factor (0 .. 1.5), sorted flag, kNN depth and queries with
    duplicated components (the last duplicate carries the value; the summary merge consumes the first only)."""
    r0 = np.random.default_rng(seed)
    build = dict(n_postings=int(r0.integers(5, 120)), centroid_fraction=float(r0.choice([0.05, 0.1, 0.2, 0.5])),
                 summary_energy=float(r0.choice([0.2, 0.4, 0.7, 1.0])), max_fraction=float(r0.choice([1.0, 1.5, 3.0])),
                 min_cluster_size=int(r0.integers(1, 4)))
    index, rng = random_index(seed, n_docs=int(r0.integers(50, 500)), dim=int(r0.integers(20, 200)), **build)
    n_knn = int(r0.integers(0, 5))
    if n_knn:
        g = rng.integers(0, index.len, size=(index.len, 5), dtype=np.uint64)
        g[rng.random(g.shape) < 0.2] = PAD
        index.set_knn(g)
    py = PyIndex(index)
    k, cut = int(r0.integers(1, 40)), int(r0.integers(1, 8))
    hf, srt = float(r0.choice([0.0, 0.5, 0.8, 0.9, 1.0, 1.5])), bool(r0.integers(0, 2))
    for _ in range(12):
        nnz = int(rng.integers(1, 10))
        c = np.sort(rng.choice(index.dim, size=nnz, replace=True)).astype(np.uint32)
        v = (rng.random(nnz) * 2).astype(np.float32)
        ids, sc, cnt, st = oracle_mod.batch_search(index.view, np.array([0, nnz], np.uint64), c, v, k, cut, hf,
                                                   n_knn=n_knn, first_sorted=srt, n_threads=1)
        p_ids, p_sc, ev = py.search(c, v, k, cut, hf, n_knn=n_knn, first_sorted=srt)
        assert cnt[0] == len(p_ids) and ids[0, : cnt[0]].tolist() == p_ids
        assert np.array_equal(sc[0, : cnt[0]], np.array(p_sc, np.float32)) and st["blocks_evaluated"] == ev


def test_dotvbyte_long_records_and_mixed_chunks():
    """DotVByte records with more than 64 chunks (several directory entries, `wide chunks before` > 0), narrow and wide
    chunks mixed, first components above and below 256, an empty document: builder, independent Python decoder and the
    library's host decoder agree on every component and code."""
    from seismic_b200 import Dataset, HostIndex
    rng = np.random.default_rng(11)
    comps, vals = [], []
    for i in range(120):
        n = int(rng.integers(1, 1400))
        if i % 3 == 0:    # dense low range: almost every gap fits one byte (narrow chunks)
            c = np.sort(rng.choice(4000, size=min(n, 3000), replace=False))
        elif i % 3 == 1:  # sparse: almost every gap needs two bytes
            c = np.sort(rng.choice(65000, size=min(n, 200), replace=False))
        else:             # clusters: runs of small gaps separated by large ones
            base = np.sort(rng.choice(60000, size=12, replace=False))
            c = np.unique(np.concatenate([b + rng.choice(300, size=min(n // 12 + 1, 250), replace=False) for b in base]))
        comps.append(c.astype(np.uint32))
        vals.append((rng.random(len(c), dtype=np.float32) * 3 + 0.01).astype(np.float32))
    comps.append(np.empty(0, np.uint32)); vals.append(np.empty(0, np.float32))
    index = HostIndex.build(Dataset.from_lists(comps, vals, dim=65536), n_postings=20)
    vb = index.convert_to_dotvbyte()
    off, dc, dv, codes = decode_dotvbyte(vb)   # asserts the directory, the wide flags and the padding on the way
    o0, c0, v0 = index.forward_csr()
    assert np.array_equal(off, o0) and np.array_equal(dc, c0)
    assert int(np.diff(o0.astype(np.int64)).max()) > 64 * 8, "no record with more than one directory entry"
    assert np.all(np.abs(dv - v0) <= float(vb.view.value_scale) / 2 + 1e-6)
    for d in (0, 1, 2, 57, len(comps) - 1):
        gc, gv = vb.get_doc(d)
        assert np.array_equal(gc, dc[int(off[d]):int(off[d + 1])]) and np.array_equal(gv, dv[int(off[d]):int(off[d + 1])])
    f16, packed = index.space_usage()["forward"], vb.space_usage()["forward"]
    assert packed < 0.85 * f16
