"""Pins the CPU oracle (and the CPU index builder it searches) against every known-answer test the
reference holds for the query hot path.  CPU only."""
import json
from pathlib import Path

import numpy as np
import pytest

from seismic_b200 import Dataset, HostIndex
from seismic_b200 import _native as N

GOLD = json.loads((Path(__file__).parent / "golden" / "reference_known_answers.json").read_text())


def _dataset(case):
    comps = [np.array(c, dtype=np.uint32) for c, _ in case["docs"]]
    vals = [np.array(v, dtype=np.float32) for _, v in case["docs"]]
    return Dataset.from_lists(comps, vals, dim=case["dim"])


def _search(oracle_mod, index, case, order):
    qc = np.array(case["query"][0], dtype=np.uint32)
    qv = np.array(case["query"][1], dtype=np.float32)
    off = np.array([0, len(qc)], dtype=np.uint64)
    ids, scores, counts, _ = oracle_mod.batch_search(
        index.view, off, qc, qv, case["k"], case["query_cut"], case["heap_factor"], case["n_knn"],
        case["first_sorted"], order=order)
    n = int(counts[0])
    return ids[0, :n].tolist(), scores[0, :n].tolist(), ids[0, n:], scores[0, n:]


@pytest.mark.parametrize("value_kind", [N.VAL_F32, N.VAL_F16, N.VAL_BF16])
@pytest.mark.parametrize("order", [0, 1])
@pytest.mark.parametrize("name", ["test_empty_vectors", "rust_usage_example"])
def test_reference_known_answers(oracle_mod, name, order, value_kind):
    case = GOLD[name]
    ds = _dataset(case)
    assert (len(ds), ds.dim, ds.nnz) == (case["len"], case["dim"], case["nnz"])
    index = HostIndex.build(ds, value_kind=value_kind)  # Configuration::default()
    assert (index.len, index.dim, index.nnz) == (case["len"], case["dim"], case["nnz"])
    ids, scores, pad_ids, pad_scores = _search(oracle_mod, index, case, order)
    assert ids == case["expected_ids"]          # best first; empty docs never retrieved; < k results allowed
    assert scores == case["expected_scores"]    # small integers: exact in every encoding
    assert (pad_ids == N.PAD_ID).all() and np.isneginf(pad_scores).all()


def test_convert_dataset_preserves_postings():
    """reference src/inverted_index.rs:774-807: every posting refers to a valid doc and some postings exist."""
    ds = Dataset.from_lists([np.array([0, 2], np.uint32), np.array([1, 3], np.uint32)],
                            [np.array([1.0, 2.0], np.float32), np.array([3.0, 4.0], np.float32)], dim=4)
    index = HostIndex.build(ds)
    a = index.arrays()
    starts = (a["postings"] >> np.uint64(16)).astype(np.int64)
    lens = (a["postings"] & np.uint64(0xFFFF)).astype(np.int64)
    assert len(starts) > 0
    doc = np.searchsorted(a["fwd_offsets"].astype(np.int64), starts, side="right") - 1
    assert ((doc >= 0) & (doc < index.len)).all()
    assert (a["fwd_offsets"][doc + 1].astype(np.int64) - a["fwd_offsets"][doc].astype(np.int64) == lens).all()


# ---- QuantizedSummary built directly from a dataset of summary vectors (quantized_summary.rs:289-406 and
# ---- utils.rs:68-90 restated in numpy), laid out as list 0 of an otherwise empty index view.
def _summary_view(vectors, dim):
    B = len(vectors)
    mins = np.zeros(B, np.float32)
    quants = np.zeros(B, np.float32)
    trip = []
    for b, (c, v) in enumerate(vectors):
        v = v.astype(np.float32)
        mn, mx = np.float32(v.min()), np.float32(v.max())
        quant = np.float32((mx - mn) / np.float32(255.0))
        with np.errstate(invalid="ignore", divide="ignore"):
            q = (v - mn) / quant
        code = np.where(np.isnan(q), 0, np.clip(np.floor(np.abs(q) + 0.5) * np.sign(q), 0, 255)).astype(np.uint8)
        mins[b], quants[b] = mn, quant
        trip += [(int(ci), b, int(co)) for ci, co in zip(c, code)]
    trip.sort(key=lambda t: (t[0], t[1]))
    sc_comp = np.array(sorted({t[0] for t in trip}), np.uint32)
    run = np.zeros(len(sc_comp) + 1, np.uint32)
    pos = {int(c): i for i, c in enumerate(sc_comp)}
    for t in trip:
        run[pos[t[0]] + 1] += 1
    run = np.cumsum(run, dtype=np.uint32)
    ent_blk = np.array([t[1] for t in trip], np.uint16)
    ent_code = np.array([t[2] for t in trip], np.uint8)

    def starts(n_first):  # list 0 owns everything, lists 1.. are empty
        a = np.full(dim + 1, n_first, np.uint64)
        a[0] = 0
        return a
    keep = dict(
        list_post_start=np.zeros(dim + 1, np.uint64), postings=np.zeros(1, np.uint64),
        list_blk_start=starts(B), blk_post_off=np.zeros(B + dim, np.uint32), blk_min=mins, blk_quant=quants,
        list_sc_start=starts(len(sc_comp)), sc_comp=sc_comp, list_ent_start=starts(len(trip)),
        sc_run_off=np.concatenate([run, np.zeros(dim - 1, np.uint32)]), ent_blk=ent_blk, ent_code=ent_code,
        fwd_offsets=np.zeros(1, np.uint64))
    view = N.IndexView()
    view.comp_bits, view.value_kind, view.n_docs, view.dim, view.value_scale = 32, N.VAL_F32, 0, dim, 1.0
    for name, arr in keep.items():
        setattr(view, name, arr.ctypes.data)
    return view, keep


def _merge_ip(qc, qv, vc, vv):
    _, ia, ib = np.intersect1d(qc, vc, assume_unique=True, return_indices=True)
    return float(np.sum(qv[ia].astype(np.float64) * vv[ib].astype(np.float64)))


def test_distances_iter(oracle_mod):
    """reference src/quantized_summary.rs:519-598: distances() == exact inner product within 1e-5 (unit values)."""
    spec = GOLD["test_distances_iter"]
    rng = np.random.default_rng(142)
    n_vecs = int(rng.integers(spec["n_vecs_range"][0], spec["n_vecs_range"][1] + 1))
    dim = int(rng.integers(spec["dim_range"][0], spec["dim_range"][1] + 1))

    def rand_vec(values_one):
        nnz = int(rng.integers(spec["nnz_range"][0], spec["nnz_range"][1] + 1))
        c = np.sort(rng.choice(dim, size=nnz, replace=False)).astype(np.uint32)
        v = np.full(nnz, spec["value"], np.float32) if values_one else rng.random(nnz, dtype=np.float32)
        return c, v
    vectors = [rand_vec(True) for _ in range(n_vecs)]
    queries = [rand_vec(False) for _ in range(spec["n_random_queries"])] + vectors
    view, keep = _summary_view(vectors, dim)
    for qc, qv in queries:
        got = oracle_mod.summary_distances(view, 0, qc, qv, n_vecs)
        want = np.array([_merge_ip(qc, qv, vc, vv) for vc, vv in vectors])
        assert np.abs(got - want).max() < spec["tolerance"]


def test_summary_quantization_roundtrip(oracle_mod):
    """Non-degenerate summaries: dequantised estimate is within quant/2 per matched component of the exact dot."""
    rng = np.random.default_rng(7)
    dim = 5000
    vectors = []
    for _ in range(40):
        c = np.sort(rng.choice(dim, size=60, replace=False)).astype(np.uint32)
        vectors.append((c, rng.random(60, dtype=np.float32) * 3))
    view, keep = _summary_view(vectors, dim)
    for _ in range(20):
        qc = np.sort(rng.choice(dim, size=200, replace=False)).astype(np.uint32)
        qv = rng.random(200, dtype=np.float32)
        got = oracle_mod.summary_distances(view, 0, qc, qv, len(vectors))
        for b, (vc, vv) in enumerate(vectors):
            matched = np.intersect1d(qc, vc, return_indices=True)[1]
            bound = 0.5 * float(keep["blk_quant"][b]) * float(qv[matched].sum()) + 1e-4
            assert abs(got[b] - _merge_ip(qc, qv, vc, vv)) <= bound


def test_oracle_rejects_bad_queries(oracle_mod, synth_small):
    _, _, index = synth_small
    off = np.array([0, 2], np.uint64)
    with pytest.raises(ValueError):   # unsorted (reference asserts, src/inverted_index.rs:172-175)
        oracle_mod.batch_search(index.view, off, np.array([5, 3], np.uint32), np.ones(2, np.float32), 10, 3, 0.8)
    with pytest.raises(ValueError):   # component >= dim
        oracle_mod.batch_search(index.view, off, np.array([5, index.dim], np.uint32), np.ones(2, np.float32), 10, 3, 0.8)
    with pytest.raises(ValueError):   # k == 0 (KHeap::new asserts)
        oracle_mod.batch_search(index.view, off, np.array([3, 5], np.uint32), np.ones(2, np.float32), 0, 3, 0.8)


def test_oracle_orders_agree_within_tolerance(oracle_mod, synth_small):
    """ORDER_LANES8 (what the kernel computes) vs ORDER_SEQ: scores within 1e-4, recall identical."""
    from seismic_b200 import recall_at_k
    _, queries, index = synth_small
    a = oracle_mod.batch_search(index.view, queries.offsets, queries.comps, queries.values, 10, 3, 0.8, order=0)
    b = oracle_mod.batch_search(index.view, queries.offsets, queries.comps, queries.values, 10, 3, 0.8, order=1)
    assert (a[2] == b[2]).all()
    same = a[0] == b[0]
    assert same.mean() > 0.999                       # near-tie flips are possible but vanishingly rare
    fin = np.isfinite(a[1]) & same
    assert np.abs(a[1][fin] - b[1][fin]).max() <= 1e-4
    ex = oracle_mod.exact_search(index.view, queries.offsets, queries.comps, queries.values, 10)
    assert abs(recall_at_k(ex[0], ex[2], a[0], a[2]) - recall_at_k(ex[0], ex[2], b[0], b[2])) < 1e-3


def test_oracle_properties(oracle_mod, synth_pruned):
    """Size-independent properties: best-first order, no duplicates, heap_factor=0 + full cut == exact search."""
    _, queries, index = synth_pruned
    ids, scores, counts, st = oracle_mod.batch_search(index.view, queries.offsets, queries.comps, queries.values,
                                                       10, 3, 0.8, first_sorted=True)
    for q in range(len(counts)):
        n = int(counts[q])
        assert (np.diff(scores[q, :n]) <= 0).all()
        assert len(set(ids[q, :n].tolist())) == n
    assert st["blocks_evaluated"] < st["blocks_total"]      # pruning does skip blocks
    # threads do not change results
    ids2, scores2, counts2, _ = oracle_mod.batch_search(index.view, queries.offsets, queries.comps, queries.values,
                                                         10, 3, 0.8, first_sorted=True, n_threads=4)
    assert (ids == ids2).all() and (scores == scores2).all() and (counts == counts2).all()
