"""Knn::refine and the kNN graph (reference src/inverted_index.rs:430-593).

CPU part: the oracle's refine against a hand-computed case and against its own definition (top-k of heap U
neighbours).  GPU part (-m gpu): the CUDA refine and the GPU graph build against the oracle."""
import numpy as np
import pytest

from seismic_b200 import Dataset, GpuIndex, HostIndex
from seismic_b200 import _native as N

PAD = np.uint64(N.PAD_ID)


def tiny_index():
    # d0 = {0: 1.0}, d1 = {1: 1.0}, d2 = {0: 0.5, 1: 0.5}, d3 = {} (empty), d4 = {2: 2.0}
    comps = [np.array([0], np.uint32), np.array([1], np.uint32), np.array([0, 1], np.uint32), np.array([], np.uint32),
             np.array([2], np.uint32)]
    vals = [np.array([1.0], np.float32), np.array([1.0], np.float32), np.array([0.5, 0.5], np.float32),
            np.array([], np.float32), np.array([2.0], np.float32)]
    return HostIndex.build(Dataset.from_lists(comps, vals, dim=3))


def csr(comps, vals):
    return np.array([0, len(comps)], np.uint64), np.array(comps, np.uint32), np.array(vals, np.float32)


def test_refine_known_answer(oracle_mod):
    """query {0: 1, 1: 1}, query_cut 1 -> only list 0 is visited (ties go to the earlier component): d0 and d2, both
    1.0.  Graph: d0 -> [d1, d4], d2 -> [d1, -], n_knn 1 -> d1 (1.0) is scored once and enters; n_knn 2 also scores d4
    (0.0, enters while the heap has room).  Equal scores: smaller forward offset first."""
    index = tiny_index()
    off, qc, qv = csr([0, 1], [1.0, 1.0])
    ids, scores, counts, _ = oracle_mod.batch_search(index.view, off, qc, qv, 4, 1, 0.8, n_knn=0, first_sorted=True)
    assert counts[0] == 2 and ids[0, :2].tolist() == [0, 2]
    # n_knn > 0 without a graph: the refine is skipped (`if n_knn > 0 && let Some(knn)`, src/inverted_index.rs:215-217)
    ids2, _, counts2, _ = oracle_mod.batch_search(index.view, off, qc, qv, 4, 1, 0.8, n_knn=1, first_sorted=True)
    assert counts2[0] == 2 and ids2[0, :2].tolist() == [0, 2]
    graph = np.array([[1, 4], [0, PAD], [1, PAD], [PAD, PAD], [PAD, PAD]], dtype=np.uint64)
    index.set_knn(graph)
    ids, scores, counts, st = oracle_mod.batch_search(index.view, off, qc, qv, 4, 1, 0.8, n_knn=1, first_sorted=True)
    assert counts[0] == 3 and ids[0, :3].tolist() == [0, 1, 2] and scores[0, :3].tolist() == [1.0, 1.0, 1.0]
    assert st["docs_scored"] == 3  # d1 is scored once although two retained documents name it
    ids, scores, counts, _ = oracle_mod.batch_search(index.view, off, qc, qv, 4, 1, 0.8, n_knn=5, first_sorted=True)
    assert counts[0] == 4 and ids[0].tolist() == [0, 1, 2, 4] and scores[0, 3] == 0.0
    ids, _, counts, _ = oracle_mod.batch_search(index.view, off, qc, qv, 2, 1, 0.8, n_knn=2, first_sorted=True)
    assert counts[0] == 2 and ids[0].tolist() == [0, 1]  # k = 2: d1 replaces d2 (same score, smaller offset)
    index.set_knn(None)
    assert index.knn is None and index.view.knn_dim == 0


def oracle_graph(oracle_mod, index, nknn):
    """Knn::new restated with the oracle: N self-searches, own id dropped, first nknn kept (PAD when short)."""
    off, comps, vals = index.forward_csr()
    ids, _, counts, _ = oracle_mod.batch_search(index.view, off, comps, vals, nknn + 1, 10, 0.7, n_knn=0,
                                                first_sorted=False)
    out = np.full((index.len, nknn), PAD, dtype=np.uint64)
    for d in range(index.len):
        row = [int(x) for x in ids[d, : counts[d]] if int(x) != d][:nknn]
        out[d, : len(row)] = row
    return out


def test_refine_equals_topk_of_heap_and_neighbours(oracle_mod, synth_small):
    """Definition check on synthetic data: refine == k best of (result without refine U their first n_knn neighbours),
    and it can only improve the k-th score."""
    _, q, index = synth_small
    rng = np.random.default_rng(5)
    graph = rng.integers(0, index.len, size=(index.len, 6), dtype=np.uint64)
    index.set_knn(graph)
    try:
        nq = 50
        off = q.offsets[: nq + 1]
        qc, qv = q.comps[: int(off[-1])], q.values[: int(off[-1])]
        base = oracle_mod.batch_search(index.view, off, qc, qv, 10, 3, 0.8, n_knn=0, first_sorted=True)
        ref = oracle_mod.batch_search(index.view, off, qc, qv, 10, 3, 0.8, n_knn=4, first_sorted=True)
        ex = oracle_mod.exact_search(index.view, off, qc, qv, index.len)  # all scores, to look candidates up
        for i in range(nq):
            score_of = {int(d): float(s) for d, s in zip(ex[0][i, : ex[2][i]], ex[1][i, : ex[2][i]])}
            cand = set(int(d) for d in base[0][i, : base[2][i]])
            for d in list(cand):
                cand.update(int(x) for x in graph[d, :4])
            cand = {d for d in cand if d in score_of}  # documents sharing no term with the query score 0: keep them too
            want = sorted(((score_of.get(d, 0.0), d) for d in cand), key=lambda t: (-t[0], t[1]))[:10]
            got = list(zip(ref[1][i, : ref[2][i]].tolist(), ref[0][i, : ref[2][i]].tolist()))
            assert [round(s, 4) for s, _ in got] == [round(s, 4) for s, _ in want][: len(got)]
            assert ref[1][i, ref[2][i] - 1] >= base[1][i, base[2][i] - 1]
    finally:
        index.set_knn(None)


def test_api_knn_roundtrip_without_gpu(tmp_path):
    """save_knn / load_knn / knn_len and the reference's error behaviour, on a host-only index."""
    from seismic_b200 import api
    idx = api.SeismicIndexRaw(tiny_index())
    assert idx.knn_len == 0
    with pytest.raises(ValueError):
        idx.save_knn(str(tmp_path / "g"))  # PyValueError: no graph attached
    graph = np.array([[1, 4], [0, PAD], [1, PAD], [PAD, PAD], [PAD, PAD]], dtype=np.uint64)
    idx._host.set_knn(graph)
    assert idx.knn_len == 2
    idx.save_knn(str(tmp_path / "g"))
    other = api.SeismicIndexRaw(tiny_index())
    other.load_knn(str(tmp_path / "g.knn.seismic"), nknn=1)
    assert other.knn_len == 1 and other._host.knn[:, 0].tolist() == graph[:, 0].tolist()
    with pytest.raises(ValueError):
        other.load_knn(str(tmp_path / "g.knn.seismic"), nknn=3)  # more than the file holds
    with pytest.raises(OSError):
        other.load_knn(str(tmp_path / "missing.knn.seismic"))


# ---------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_refine_known_answer(oracle_mod):
    index = tiny_index()
    graph = np.array([[1, 4], [0, PAD], [1, PAD], [PAD, PAD], [PAD, PAD]], dtype=np.uint64)
    index.set_knn(graph)
    gpu = GpuIndex(index, 0)  # the view carries the graph
    off, qc, qv = csr([0, 1], [1.0, 1.0])
    for k, n_knn in ((4, 1), (4, 5), (2, 2), (4, 0)):
        ref = oracle_mod.batch_search(index.view, off, qc, qv, k, 1, 0.8, n_knn=n_knn, first_sorted=True)
        got = gpu.batch_search(off, qc, qv, k, 1, 0.8, n_knn=n_knn, first_sorted=True)
        assert (got[2] == ref[2]).all() and (got[0] == ref[0]).all() and np.array_equal(got[1], ref[1]), (k, n_knn)
    gpu.set_knn(None)  # no graph: n_knn is ignored, as in the reference
    got = gpu.batch_search(off, qc, qv, 4, 1, 0.8, n_knn=1)
    assert got[2][0] == 2 and got[0][0, :2].tolist() == [0, 2]


@pytest.mark.gpu
@pytest.mark.parametrize("k,n_knn,dim_knn", [(10, 5, 8), (10, 8, 8), (100, 3, 4), (1, 2, 2), (10, 40, 40)])
def test_gpu_refine_parity(oracle_mod, synth_pruned, k, n_knn, dim_knn):
    """Random graph (duplicates, self loops, PAD entries, empty documents as neighbours): ids, score bits and counts
    of the CUDA refine equal the oracle's; register heap (k <= 32) and shared-memory heap (k = 100)."""
    _, q, index = synth_pruned
    rng = np.random.default_rng(k * 131 + n_knn)
    graph = rng.integers(0, index.len, size=(index.len, dim_knn), dtype=np.uint64)
    graph[rng.random(graph.shape) < 0.05] = PAD
    index.set_knn(graph)
    try:
        gpu = GpuIndex(index, 0)
        ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, 3, 0.8, n_knn=n_knn, first_sorted=True)
        got = gpu.batch_search(q.offsets, q.comps, q.values, k, 3, 0.8, n_knn=n_knn, first_sorted=True)
        assert (got[2] == ref[2]).all()
        assert (got[0] == ref[0]).all()
        assert np.array_equal(got[1].view(np.uint32), ref[1].view(np.uint32))
    finally:
        index.set_knn(None)


@pytest.mark.gpu
def test_gpu_build_knn_matches_oracle_graph(oracle_mod):
    """Knn::new as GPU batches == the same N self-searches on the oracle; then search with the built graph."""
    from seismic_b200 import api
    cfg = Dataset.synth_config(4000, dim=3000)
    docs = Dataset.synth_documents(cfg)
    host = HostIndex.build(docs, n_postings=400, centroid_fraction=0.2)
    idx = api.SeismicIndexRaw(host)
    idx.build_knn(5, batch_docs=1500)
    want = oracle_graph(oracle_mod, host, 5)
    assert idx.knn_len == 5
    assert np.array_equal(host.knn, want)
    q = Dataset.synth_queries(cfg, 100)
    ref = oracle_mod.batch_search(host.view, q.offsets, q.comps, q.values, 10, 2, 0.9, n_knn=5, first_sorted=True)
    got = idx.gpu.batch_search(q.offsets, q.comps, q.values, 10, 2, 0.9, n_knn=5, first_sorted=True)
    assert (got[0] == ref[0]).all() and np.array_equal(got[1], ref[1])


@pytest.mark.gpu
def test_gpu_refine_many_queries_per_cta(oracle_mod):
    """Regression: with several queries per CTA and query_cut 1, a selection pass that finds nothing leaves the list
    loop without a barrier; the refine must not reuse shared state a slower warp is still reading (found with
    compute-sanitizer --tool racecheck)."""
    cfg = Dataset.synth_config(30000, dim=2000)
    index = HostIndex.build(Dataset.synth_documents(cfg), n_postings=600, centroid_fraction=0.2)
    q = Dataset.synth_queries(cfg, 6000)
    rng = np.random.default_rng(3)
    index.set_knn(rng.integers(0, index.len, size=(index.len, 10), dtype=np.uint64))
    gpu = GpuIndex(index, 0)
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, 10, 1, 0.9, n_knn=10, first_sorted=True)
    for _ in range(3):
        got = gpu.batch_search(q.offsets, q.comps, q.values, 10, 1, 0.9, n_knn=10, first_sorted=True)
        assert (got[2] == ref[2]).all() and (got[0] == ref[0]).all()
        assert np.array_equal(got[1].view(np.uint32), ref[1].view(np.uint32))
