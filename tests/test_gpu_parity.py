"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same index and
queries.  Bar: identical doc ids, bit-identical f32 scores (the kernel and the oracle use the same summation
order), identical counts.  Run with `pytest -m gpu` on the B200 box."""
import json
from pathlib import Path

import numpy as np
import pytest

from seismic_b200 import Dataset, GpuIndex, HostIndex, recall_at_k
from seismic_b200 import _native as N

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).parent / "golden" / "reference_known_answers.json").read_text())


def assert_same(gpu, ref, what=""):
    gi, gs, gc = gpu[:3]
    ri, rs, rc = ref[:3]
    assert (gc == rc).all(), f"{what}: counts differ for {int((gc != rc).sum())} queries"
    bad = np.nonzero((gi != ri).any(axis=1))[0]
    assert len(bad) == 0, f"{what}: ids differ for {len(bad)} queries, first {bad[:5]}: {gi[bad[0]]} vs {ri[bad[0]]}"
    assert (gs.view(np.uint32) == rs.view(np.uint32)).all() or np.array_equal(gs, rs), f"{what}: scores differ"


@pytest.fixture(scope="module")
def gpu_small(synth_small):
    return GpuIndex(synth_small[2], 0)


@pytest.fixture(scope="module")
def gpu_pruned(synth_pruned):
    return GpuIndex(synth_pruned[2], 0)


@pytest.mark.parametrize("k,cut,hf,srt", [
    (10, 3, 0.8, True), (10, 3, 0.8, False), (10, 1, 0.9, True), (10, 10, 1.0, True), (1, 3, 0.8, True),
    (100, 5, 0.7, True), (33, 4, 0.0, False), (10, 1000, 0.8, True)])
def test_parity_small(oracle_mod, synth_small, gpu_small, k, cut, hf, srt):
    _, q, index = synth_small
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    got = gpu_small.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    assert_same(got, ref, f"k={k} cut={cut} hf={hf} sorted={srt}")
    assert gpu_small.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]


@pytest.mark.parametrize("k,cut,hf,srt", [(10, 3, 0.8, True), (10, 3, 0.8, False), (100, 8, 0.9, True), (10, 3, 1.2, True)])
def test_parity_pruned(oracle_mod, synth_pruned, gpu_pruned, k, cut, hf, srt):
    _, q, index = synth_pruned
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    got = gpu_pruned.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    assert_same(got, ref, f"k={k} cut={cut} hf={hf} sorted={srt}")
    assert gpu_pruned.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]
    assert gpu_pruned.last_stats["docs_scored"] >= ref[3]["docs_scored"]


@pytest.mark.parametrize("hq", [3, 1, 2, 0])
@pytest.mark.parametrize("wave,first,bucket", [(1, 1, 1), (64, 8, 0), (64, 8, 1), (512, 512, 1), (4096, 4096, 0), (4096, 4096, 1)])
def test_wave_sizes_do_not_change_results(oracle_mod, synth_pruned, wave, first, hq, bucket):
    """The speculative wave scheduler is a performance knob only: any wave size replays to the same heap,
    in the hash-query kernel (hq=1) as well as in the dense-query kernel (hq=0)."""
    _, q, index = synth_pruned
    g = GpuIndex(index, 0)
    g.set_option("hq", hq)
    g.set_option("bucket", bucket)
    g.set_option("wave_docs", wave)
    g.set_option("first_wave_docs", first)
    g.set_option("hq_wave_docs", min(wave, 1024))     # blocks larger than the wave buffer are split
    g.set_option("hq_first_wave_docs", min(first, 1024))
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, 10, 3, 0.8, first_sorted=True)
    got = g.batch_search(q.offsets, q.comps, q.values, 10, 3, 0.8, first_sorted=True)
    assert_same(got, ref, f"wave={wave}")


@pytest.mark.parametrize("hq", [3, 1, 2, 0])
def test_dense_and_hash_kernels_agree(oracle_mod, synth_small, hq):
    _, q, index = synth_small
    g = GpuIndex(index, 0)
    g.set_option("hq", hq)
    for k, cut, hf, srt in [(10, 3, 0.8, True), (100, 6, 0.9, False)]:
        ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        assert_same(got, ref, f"hq={hq} k={k}")


def test_lists_with_thousands_of_blocks(oracle_mod):
    """centroid_fraction 0.5 gives lists with > 1024 blocks: k_est accumulates in global memory instead of shared
    memory, k_order sorts up to 4096 keys, and selection needs several rounds of 256 positions per list."""
    from conftest import build_synth
    _, q, index = build_synth(20000, 200, dim=300, n_postings=3000, centroid_fraction=0.5, max_fraction=2.0,
                              min_cluster_size=0)
    a = index.arrays()
    assert int(np.diff(a["list_blk_start"]).max()) > 1024
    g = GpuIndex(index, 0)
    for srt in (True, False):
        ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, 10, 4, 0.9, first_sorted=srt)
        got = g.batch_search(q.offsets, q.comps, q.values, 10, 4, 0.9, first_sorted=srt)
        assert_same(got, ref, f"many blocks sorted={srt}")


@pytest.fixture(scope="module")
def synth_lv():
    """Large-vocabulary index: 120 k vocabulary, u32 components (SeismicIndexLV)."""
    from conftest import build_synth
    return build_synth(30000, 300, dim=120000, comp_bits=32, n_postings=200, centroid_fraction=0.2)


@pytest.mark.parametrize("k,cut,hf,srt", [(10, 3, 0.8, True), (100, 5, 0.9, True), (100, 8, 0.8, False), (7, 2, 1.0, True)])
def test_parity_large_vocabulary(oracle_mod, synth_lv, k, cut, hf, srt):
    _, q, index = synth_lv
    assert index.comp_bits == 32 and index.dim == 120000
    g = GpuIndex(index, 0)
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    assert_same(got, ref, f"LV k={k} cut={cut} hf={hf} sorted={srt}")
    assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]


def _long_queries(docs, dim, rng, n=12):
    """Queries with more than 255 components: whole documents glued together + dense ramps (the document-as-query
    searches of Knn::new, src/inverted_index.rs:448-500, are of this kind on real corpora)."""
    comps, vals = [], []
    for i in range(n):
        c = np.unique(np.concatenate([docs.vector(j)[0] for j in rng.integers(0, len(docs), size=4 + i)]))
        if i % 3 == 0:
            c = np.unique(np.concatenate([c, np.arange(i, dim, max(1, dim // 700), dtype=np.uint32)]))
        comps.append(c.astype(np.uint32))
        vals.append((rng.random(len(c), dtype=np.float32) * 2 + 0.01).astype(np.float32))
    comps.append(np.array([3, 3, 9], np.uint32)); vals.append(np.array([1.0, 2.0, 0.5], np.float32))  # short, with a duplicate
    off = np.zeros(len(comps) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in comps])
    assert max(len(c) for c in comps) > 255
    return off, np.concatenate(comps), np.concatenate(vals)


def test_large_vocabulary_long_queries(oracle_mod, synth_lv):
    """The reference accepts any sorted query (src/inverted_index.rs:172-175).  u32 indexes have no dense-query kernel:
    queries with more than 255 components take the sorted-query kernel in a second pass."""
    docs, q, index = synth_lv
    g = GpuIndex(index, 0)
    off, qc, qv = _long_queries(docs, index.dim, np.random.default_rng(11))
    for k, cut, hf, srt in [(10, 3, 0.8, True), (100, 12, 0.9, False), (10, 1000, 0.7, True)]:
        ref = oracle_mod.batch_search(index.view, off, qc, qv, k, cut, hf, first_sorted=srt)
        got = g.batch_search(off, qc, qv, k, cut, hf, first_sorted=srt)
        assert_same(got, ref, f"LV long queries k={k} cut={cut}")
    # mixed with ordinary queries in one batch, device entry order preserved
    off2 = np.concatenate([q.offsets, q.offsets[-1] + off[1:]]).astype(np.uint64)
    qc2, qv2 = np.concatenate([q.comps, qc]), np.concatenate([q.values, qv])
    ref = oracle_mod.batch_search(index.view, off2, qc2, qv2, 10, 3, 0.8)
    got = g.batch_search(off2, qc2, qv2, 10, 3, 0.8)
    assert_same(got, ref, "LV mixed batch")


@pytest.mark.parametrize("value_kind", [N.VAL_BF16, N.VAL_F32, N.VAL_FIXEDU8, N.VAL_FIXEDU16])
def test_parity_large_vocabulary_value_encodings(oracle_mod, value_kind):
    """u32 components x bf16 / f32 / fixedu8 / fixedu16 (the other half of the reference's encoding matrix,
    src/bin/perf_inverted_index.rs:95-139)."""
    from conftest import build_synth
    docs, q, index = build_synth(20000, 200, dim=90000, comp_bits=32, n_postings=150, centroid_fraction=0.2,
                                 value_kind=value_kind)
    assert index.value_kind == value_kind and index.comp_bits == 32
    g = GpuIndex(index, 0)
    for k, cut, hf, srt in [(10, 3, 0.8, True), (50, 6, 0.9, False)]:
        ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        assert_same(got, ref, f"u32 value_kind={value_kind} k={k}")
        assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]
    off, qc, qv = _long_queries(docs, index.dim, np.random.default_rng(3), n=5)
    ref = oracle_mod.batch_search(index.view, off, qc, qv, 10, 5, 0.8)
    got = g.batch_search(off, qc, qv, 10, 5, 0.8)
    assert_same(got, ref, f"u32 value_kind={value_kind} long queries")


@pytest.mark.parametrize("k,cut,hf,srt", [(10, 3, 0.8, True), (10, 3, 0.8, False), (100, 6, 0.9, True)])
def test_parity_dotvbyte(oracle_mod, synth_pruned, k, cut, hf, srt):
    """SeismicIndexDotVByte: variable-byte gaps + u8 values decoded inside the scoring kernel."""
    _, q, index = synth_pruned
    vb = index.convert_to_dotvbyte()
    assert vb.value_kind == N.VAL_DOTVBYTE and vb.space_usage()["forward"] < 0.7 * index.space_usage()["forward"]
    g = GpuIndex(vb, 0)
    ref = oracle_mod.batch_search(vb.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    assert_same(got, ref, f"dotvbyte k={k} cut={cut} hf={hf} sorted={srt}")
    assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]
    # the u8 re-quantisation keeps the ranking close to the f16 index it was converted from
    ref16 = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    assert recall_at_k(ref16[0], ref16[2], got[0], got[2]) > 0.9


def test_dotvbyte_wide_documents(oracle_mod):
    """Documents with up to 700 components (two super-rounds of 64 chunks: the wide-chunk rank spans directory entries
    and both mask words), narrow and wide chunks mixed, an empty document, and queries of > 255 components."""
    rng = np.random.default_rng(5)
    comps, vals = [], []
    for i in range(600):
        n = int(rng.integers(1, 700))
        comps.append(np.sort(rng.choice(60000, size=n, replace=False)).astype(np.uint32))
        vals.append((rng.random(n, dtype=np.float32) * 3 + 0.01).astype(np.float32))
    comps.append(np.empty(0, np.uint32)); vals.append(np.empty(0, np.float32))        # an empty document
    index = HostIndex.build(Dataset.from_lists(comps, vals, dim=60000), n_postings=50)
    vb = index.convert_to_dotvbyte()
    g = GpuIndex(vb, 0)
    qc = [np.sort(rng.choice(60000, size=200, replace=False)).astype(np.uint32) for _ in range(40)] + [comps[3], comps[77], comps[5][:300]]   # > 255 components: sorted-query pass
    qv = [rng.random(len(c), dtype=np.float32) for c in qc]
    off = np.zeros(len(qc) + 1, np.uint64); off[1:] = np.cumsum([len(c) for c in qc])
    ref = oracle_mod.batch_search(vb.view, off, np.concatenate(qc), np.concatenate(qv), 10, 20, 0.0, first_sorted=False)
    got = g.batch_search(off, np.concatenate(qc), np.concatenate(qv), 10, 20, 0.0, first_sorted=False)
    assert_same(got, ref, "dotvbyte wide docs")


@pytest.mark.parametrize("value_kind", [N.VAL_BF16, N.VAL_F32, N.VAL_FIXEDU8, N.VAL_FIXEDU16])
def test_parity_value_encodings(oracle_mod, value_kind):
    """The other forward-index value encodings of the reference: bf16, f32, fixedu8, fixedu16."""
    from conftest import build_synth
    docs, q, index = build_synth(20000, 200, dim=3000, n_postings=500, centroid_fraction=0.15, value_kind=value_kind)
    assert index.value_kind == value_kind
    g = GpuIndex(index, 0)
    for k, cut, hf, srt in [(10, 3, 0.8, True), (50, 6, 0.9, False)]:
        ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        assert_same(got, ref, f"value_kind={value_kind} k={k}")
        assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]
    off, qc, qv = _long_queries(docs, index.dim, np.random.default_rng(4), n=5)
    ref = oracle_mod.batch_search(index.view, off, qc, qv, 10, 5, 0.8)
    got = g.batch_search(off, qc, qv, 10, 5, 0.8)
    assert_same(got, ref, f"value_kind={value_kind} long queries")


def test_small_scratch_chunks_the_batch(oracle_mod, synth_small):
    _, q, index = synth_small
    g = GpuIndex(index, 0)
    g.set_option("scratch_mb", 1)
    g.set_option("ctas", 7)
    g.set_option("hq_ctas_per_sm", 1)
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, 10, 3, 0.8)
    got = g.batch_search(q.offsets, q.comps, q.values, 10, 3, 0.8)
    assert_same(got, ref, "chunked")


@pytest.mark.parametrize("name", ["test_empty_vectors", "rust_usage_example"])
def test_reference_known_answers_on_gpu(name):
    case = GOLD[name]
    comps = [np.array(c, dtype=np.uint32) for c, _ in case["docs"]]
    vals = [np.array(v, dtype=np.float32) for _, v in case["docs"]]
    index = HostIndex.build(Dataset.from_lists(comps, vals, dim=case["dim"]))
    g = GpuIndex(index, 0)
    qc = np.array(case["query"][0], np.uint32)
    qv = np.array(case["query"][1], np.float32)
    ids, scores, counts = g.batch_search(np.array([0, len(qc)], np.uint64), qc, qv, case["k"], case["query_cut"],
                                         case["heap_factor"], first_sorted=case["first_sorted"])
    n = int(counts[0])
    assert ids[0, :n].tolist() == case["expected_ids"]
    assert scores[0, :n].tolist() == case["expected_scores"]
    assert (ids[0, n:] == N.PAD_ID).all() and np.isneginf(scores[0, n:]).all()


def test_edge_queries(oracle_mod, synth_pruned, gpu_pruned):
    """Empty query, single-term query, repeated components, a query made of one whole document, ragged batch."""
    docs, q, index = synth_pruned
    dc, dv = docs.vector(17)
    comps = [np.empty(0, np.uint32), np.array([5], np.uint32), np.array([7, 7, 9, 9, 9, 400], np.uint32),
             dc.copy(), np.arange(0, 2000, 7, dtype=np.uint32)]
    vals = [np.empty(0, np.float32), np.array([1.5], np.float32), np.array([0.5, 2.0, 1.0, 3.0, 0.1, 0.7], np.float32),
            dv.copy(), np.linspace(0.01, 2.0, len(comps[4]), dtype=np.float32)]
    off = np.zeros(len(comps) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in comps])
    qc, qv = np.concatenate(comps), np.concatenate(vals)
    for srt in (True, False):
        ref = oracle_mod.batch_search(index.view, off, qc, qv, 10, 3, 0.8, first_sorted=srt)
        got = gpu_pruned.batch_search(off, qc, qv, 10, 3, 0.8, first_sorted=srt)
        assert_same(got, ref, f"edge sorted={srt}")
        assert got[2][0] == 0 and (got[0][0] == N.PAD_ID).all()
        assert got[0][3, 0] == 17        # a document used as query retrieves itself first


def test_empty_batch_and_errors(synth_pruned, gpu_pruned):
    _, q, index = synth_pruned
    ids, scores, counts = gpu_pruned.batch_search(np.zeros(1, np.uint64), np.empty(0, np.uint32), np.empty(0, np.float32), 10, 3, 0.8)
    assert ids.shape == (0, 10) and counts.shape == (0,)
    off = np.array([0, 2], np.uint64)
    with pytest.raises(ValueError):
        gpu_pruned.batch_search(off, np.array([9, 3], np.uint32), np.ones(2, np.float32), 10, 3, 0.8)
    with pytest.raises(ValueError):
        gpu_pruned.batch_search(off, np.array([3, index.dim], np.uint32), np.ones(2, np.float32), 10, 3, 0.8)
    with pytest.raises(ValueError):
        gpu_pruned.batch_search(off, np.array([3, 9], np.uint32), np.ones(2, np.float32), 0, 3, 0.8)
    # n_knn > 0 without a kNN graph is not an error: the reference skips the refine (src/inverted_index.rs:215-217)
    gpu_pruned.batch_search(off, np.array([3, 9], np.uint32), np.ones(2, np.float32), 10, 3, 0.8, n_knn=5)
    # the index stays usable after an error
    got = gpu_pruned.batch_search(q.offsets, q.comps, q.values, 10, 3, 0.8)
    assert got[2].max() == 10


def test_exact_search_matches_oracle(oracle_mod, synth_pruned, gpu_pruned):
    _, q, index = synth_pruned
    n = 64
    off = q.offsets[: n + 1].copy()
    qc, qv = q.comps[: int(off[-1])], q.values[: int(off[-1])]
    ref = oracle_mod.exact_search(index.view, off, qc, qv, 10)
    got = gpu_pruned.exact_search(off, qc, qv, 10)
    assert_same(got, ref, "exact")
    approx = gpu_pruned.batch_search(off, qc, qv, 10, 3, 0.8)
    r = recall_at_k(got[0], got[2], approx[0], approx[2])
    assert 0.5 < r <= 1.0


def test_device_pointer_entry(oracle_mod, synth_pruned, gpu_pruned):
    """sgpu_batch_search_device with torch-owned device buffers (what the NCCL gather path uses)."""
    import torch
    _, q, index = synth_pruned
    dev = torch.device("cuda:0")
    d_off = torch.from_numpy(q.offsets.astype(np.int64)).to(dev)
    d_c = torch.from_numpy(q.comps.astype(np.int32)).to(dev)
    d_v = torch.from_numpy(q.values.copy()).to(dev)
    nq, k = len(q), 10
    d_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    d_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    st = gpu_pruned.batch_search_device(d_off.data_ptr(), d_c.data_ptr(), d_v.data_ptr(), nq, k, 3, 0.8,
                                        d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr())
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, 3, 0.8)
    got = (d_ids.cpu().numpy().view(np.uint64), d_sc.cpu().numpy(), d_cnt.cpu().numpy().view(np.uint32))
    assert_same(got, ref, "device entry")
    assert st["n_launches"] >= 5 and st["ms_search"] > 0


def test_full_scan_equals_exact(synth_pruned, gpu_pruned):
    """Property: with heap_factor = 0 nothing is skipped, so searching ALL query terms returns the exact top-k
    among documents that share a pruned posting with the query — and every returned score is the true dot."""
    docs, q, index = synth_pruned
    n = 32
    off = q.offsets[: n + 1].copy()
    qc, qv = q.comps[: int(off[-1])], q.values[: int(off[-1])]
    ids, scores, counts = gpu_pruned.batch_search(off, qc, qv, 10, 10000, 0.0, first_sorted=False)
    for i in range(n):
        dense = np.zeros(index.dim, np.float32)
        dense[qc[int(off[i]):int(off[i + 1])]] = qv[int(off[i]):int(off[i + 1])]
        for r in range(int(counts[i])):
            c, v = index.get_doc(int(ids[i, r]))
            assert abs(float(np.dot(dense[c].astype(np.float64), v.astype(np.float64))) - scores[i, r]) <= 1e-4 * max(1.0, abs(scores[i, r]))


@pytest.mark.parametrize("comp_bits,value_kind,dim", [(32, N.VAL_F16, 120000), (32, N.VAL_FIXEDU8, 90000),
                                                     (16, N.VAL_BF16, 3000), (16, N.VAL_F16, 60000)])
def test_exact_search_other_layouts(oracle_mod, comp_bits, value_kind, dim):
    """Brute-force top-k (FlatIndex stand-in, src/inverted_index_wrapper.rs:721-742) on u32 / non-f16 layouts and on a
    u16 vocabulary whose dense query does not fit shared memory (sorted-query kernel)."""
    from conftest import build_synth
    _, q, index = build_synth(12000, 40, dim=dim, comp_bits=comp_bits, n_postings=100, value_kind=value_kind)
    g = GpuIndex(index, 0)
    ref = oracle_mod.exact_search(index.view, q.offsets, q.comps, q.values, 10)
    got = g.exact_search(q.offsets, q.comps, q.values, 10)
    assert_same(got, ref, f"exact comp_bits={comp_bits} value_kind={value_kind}")


def test_group_entry_single_and_multi_device(oracle_mod, synth_pruned):
    """sgpu_group_batch_search: the batch split over the devices of the box + one NCCL gather returns exactly the
    single-device (= oracle) results, in input order.  Runs on however many GPUs are visible (1 GPU: the split and the
    gather degenerate, the entry point is still exercised)."""
    import torch
    from seismic_b200 import GpuGroup
    _, q, index = synth_pruned
    n_dev = min(torch.cuda.device_count(), 8)
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, 10, 3, 0.8)
    for devs in ([0], list(range(n_dev))):
        g = GpuGroup(index, devs)
        assert len(g) == len(devs)
        got = g.batch_search(q.offsets, q.comps, q.values, 10, 3, 0.8)
        assert_same(got, ref, f"group over {devs}")
        # a batch smaller than the group: some devices get nothing
        small = (q.offsets[:2].copy(), q.comps[: int(q.offsets[1])], q.values[: int(q.offsets[1])])
        got1 = g.batch_search(*small, 10, 3, 0.8)
        assert (got1[0][0] == ref[0][0]).all()
        with pytest.raises(ValueError):
            g.batch_search(np.array([0, 2], np.uint64), np.array([9, 3], np.uint32), np.ones(2, np.float32), 10, 3, 0.8)
        del g


@pytest.mark.parametrize("k,cut,hf,srt,wave", [(10, 3, 0.8, True, 768), (10, 5, 0.9, False, 768), (100, 4, 0.7, True, 64),
                                               (1, 1000, 0.0, False, 1)])
def test_tma_staged_records(oracle_mod, synth_pruned, synth_small, k, cut, hf, srt, wave):
    """k_search<..., TMA = true>: the records are staged round by round into shared memory by cp.async.bulk + mbarrier
    instead of per-lane 128-bit loads; results are identical (ids, score bits, evaluated blocks)."""
    for docs, q, index in (synth_pruned, synth_small):
        g = GpuIndex(index, 0)
        g.set_option("tma", 1)
        g.set_option("hq_wave_docs", wave)
        g.set_option("hq_first_wave_docs", min(wave, 128))
        ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        assert_same(got, ref, f"tma k={k} cut={cut}")
        assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"] and g.last_stats["ctas_per_sm"] >= 1


@pytest.mark.parametrize("frac,lo,hi", [(0.03, 129, 256), (0.1, 513, 1024), (0.25, 1025, 2048), (0.5, 2049, 4096)])
def test_first_list_order_size_classes(oracle_mod, frac, lo, hi):
    """sort_and_search order of the first list (src/posting_list.rs:162-166): k_order_warp sorts lists of <= 512 and of
    513..1024 blocks in registers (4 / 8 / 16 / 32 composites per lane), k_order the longer ones in shared memory — and
    the old CTA-wide kernel (`order_warp = 0`) must give the same results on all of them.  The four corpora have lists
    of up to 180, 599, 1493 and 2968 blocks (medians 84, 280, 701, 1396), so every size class is exercised."""
    from conftest import build_synth
    _, q, index = build_synth(20000, 200, dim=300, n_postings=3000, centroid_fraction=frac, max_fraction=2.0,
                              min_cluster_size=0)
    a = index.arrays()
    mx = int(np.diff(a["list_blk_start"]).max())
    assert lo <= mx <= hi, (frac, mx)
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, 10, 4, 0.9, first_sorted=True)
    g = GpuIndex(index, 0)
    for ow in (1, 0):
        g.set_option("order_warp", ow)
        got = g.batch_search(q.offsets, q.comps, q.values, 10, 4, 0.9, first_sorted=True)
        assert_same(got, ref, f"order_warp={ow} max blocks {mx}")
        assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]


@pytest.mark.parametrize("k,cut,hf,srt", [(33, 3, 0.8, True), (100, 5, 0.9, True), (128, 4, 0.7, False), (64, 1000, 0.0, True)])
def test_register_heap_for_k_up_to_128(oracle_mod, synth_pruned, synth_small, k, cut, hf, srt):
    """KHeap for 32 < k <= 128 (src/utils.rs:12-66): WideHeap (unsorted, in registers, worst tracked by warp reductions)
    against the oracle and against the sorted shared-memory heap (`wide_heap = 0`)."""
    for docs, q, index in (synth_pruned, synth_small):
        g = GpuIndex(index, 0)
        ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        for wide in (1, 0):
            g.set_option("wide_heap", wide)
            got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
            assert_same(got, ref, f"wide_heap={wide} k={k} cut={cut}")
            assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]


@pytest.mark.parametrize("var", [41, 51])
@pytest.mark.parametrize("k,cut,hf,srt", [(10, 3, 0.8, True), (100, 5, 0.9, False), (1, 1000, 0.0, True)])
def test_one_document_per_group_builds(oracle_mod, synth_pruned, k, cut, hf, srt, var):
    """k_search<256, OCC, D = 1, ...>: one document in flight per 8-lane group (4 or 5 CTAs per SM) — same results."""
    _, q, index = synth_pruned
    g = GpuIndex(index, 0)
    g.set_option("occ16", var)
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
    assert_same(got, ref, f"occ16={var} k={k} cut={cut}")
    assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]


@pytest.mark.parametrize("occ", [4, 41, 51, 3, 2])
def test_large_vocabulary_kernel_builds(oracle_mod, synth_lv, occ):
    """The u32-component kernel under its register budgets (`occ32`): identical results."""
    _, q, index = synth_lv
    g = GpuIndex(index, 0)
    g.set_option("occ32", occ)
    for k, cut, hf, srt in ((100, 5, 0.9, True), (10, 3, 0.8, False)):
        ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        got = g.batch_search(q.offsets, q.comps, q.values, k, cut, hf, first_sorted=srt)
        assert_same(got, ref, f"occ32={occ} k={k}")
        assert g.last_stats["blocks_pushed"] == ref[3]["blocks_evaluated"]


def test_page_locked_caller_buffers(oracle_mod, synth_pruned, gpu_pruned):
    """sgpu_batch_search reads page-locked query buffers and writes page-locked result buffers by DMA (no staging copy);
    pageable buffers take the staged path — same results either way, also when the two kinds are mixed."""
    from seismic_b200 import pinned_array
    _, q, index = synth_pruned
    nq, k = len(q.offsets) - 1, 10
    ref = oracle_mod.batch_search(index.view, q.offsets, q.comps, q.values, k, 3, 0.8, first_sorted=True)
    p_off, p_c, p_v = pinned_array(q.offsets.shape, np.uint64), pinned_array(q.comps.shape, np.uint32), pinned_array(q.values.shape, np.float32)
    p_off[:], p_c[:], p_v[:] = q.offsets, q.comps, q.values
    out = (pinned_array((nq, k), np.uint64), pinned_array((nq, k), np.float32), pinned_array(nq, np.uint32))
    got = gpu_pruned.batch_search(p_off, p_c, p_v, k, 3, 0.8, first_sorted=True, out=out)
    assert got[0] is out[0] and got[2] is out[2]
    assert_same(got, ref, "pinned in, pinned out")
    assert_same(gpu_pruned.batch_search(p_off, q.comps, p_v, k, 3, 0.8, first_sorted=True), ref, "mixed in, pageable out")
    out2 = (np.empty((nq, k), np.uint64), np.empty((nq, k), np.float32), np.empty(nq, np.uint32))
    assert_same(gpu_pruned.batch_search(q.offsets, q.comps, q.values, k, 3, 0.8, first_sorted=True, out=out2), ref, "pageable out=")
    with pytest.raises(ValueError):
        gpu_pruned.batch_search(q.offsets, q.comps, q.values, k, 3, 0.8, out=(out2[0][:, :5], out2[1], out2[2]))
