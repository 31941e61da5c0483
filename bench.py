#!/usr/bin/env python
"""bench.py — queries/sec of the Seismic query hot path on B200 (BASELINE.json metric), every BASELINE config.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--configs 2,r97,3,4,5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path (SeismicIndex.batch_search, Python default sorted=True) over one batch of
`--queries` synthetic queries.  Rank 0 prints ONE JSON line; its top level is BASELINE config 2 (configs[1]):
the synthetic SPLADE-v3-shaped corpus (8.8 M docs, vocab 30 522, ~120 nnz/doc, ~40 nnz/query, k=10, query_cut=3,
heap_factor=0.8; index built with the reference's Python defaults), index replicated on every GPU, 10 k queries
per GPU per step (weak scaling), results gathered on rank 0 with one NCCL gather per step.

  value      whole-job queries/s, inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e        the same through the host-buffer C-ABI call (H2D of the queries + D2H of the results inside)
  roofline / cpu_baseline / clocks as DESIGN.md §Measurement describes
  configs    the other BASELINE rows, each timed the same way (device events + host-buffer e2e) with full-batch parity
             against the CPU oracle:
               r97       the metric's operating point: first (query_cut, heap_factor) candidate with recall@10 >= 0.97
               3_strong  configs[2]: ONE batch of 10 k queries split over the N GPUs, one NCCL gather, all 10 k results
                         verified on rank 0 (strong scaling)
               4_dotvbyte  configs[3]: the same index over a DotVByte forward index (decode fused into the kernel)
               5_lv      configs[4] shape: u32 components, vocab 200 k, ~150 nnz/doc, k=100 — `--lv-docs` documents
                         (the 50 M-document index does not build within a bench run)
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--configs", default=os.environ.get("SEISMIC_BENCH_CONFIGS", "2,r97,3,4,5"),
                    help="which BASELINE rows to run (2 always runs)")
    ap.add_argument("--docs", type=int, default=int(os.environ.get("SEISMIC_BENCH_DOCS", 8_800_000)))
    ap.add_argument("--dim", type=int, default=30522)
    ap.add_argument("--queries", type=int, default=int(os.environ.get("SEISMIC_BENCH_QUERIES", 10_000)),
                    help="queries per GPU per step (weak) / per step in total (3_strong)")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--query-cut", type=int, default=3)
    ap.add_argument("--heap-factor", type=float, default=0.8)
    ap.add_argument("--sorted", type=int, default=1)
    ap.add_argument("--n-postings", type=int, default=3500)
    ap.add_argument("--centroid-fraction", type=float, default=0.1)
    ap.add_argument("--summary-energy", type=float, default=0.4)
    ap.add_argument("--max-fraction", type=float, default=1.5)
    ap.add_argument("--recall-queries", type=int, default=1000, help="queries used for recall@k vs exact (0 = skip)")
    ap.add_argument("--cpu-sample", type=int, default=2000, help="queries of the single-thread CPU sample")
    ap.add_argument("--r97", default="5:0.9,6:0.9,5:0.8,8:0.9",
                    help="query_cut:heap_factor candidates, cheapest first; the first with recall@10 >= 0.97 is timed")
    ap.add_argument("--lv-docs", type=int, default=int(os.environ.get("SEISMIC_BENCH_LV_DOCS", 6_000_000)))
    ap.add_argument("--lv-dim", type=int, default=200_000)
    ap.add_argument("--lv-k", type=int, default=100)
    ap.add_argument("--lv-nnz", type=float, default=145.0, help="median doc nnz of the LV corpus (mean ~150)")
    ap.add_argument("--keep-index", action="store_true")
    return ap.parse_args()


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# Everything except the final JSON line goes to stderr: libraries (e.g. NCCL's version banner) write to fd 1.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj) -> None:
    _REAL_STDOUT.write(json.dumps(obj) + "\n")
    _REAL_STDOUT.flush()


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (profiling recipe's clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def workload_name(a) -> str:
    return ("synthetic SPLADE-v3-shaped: %s docs, vocab %d, ~120 nnz/doc, ~40 nnz/query, k=%d, query_cut=%d, "
            "heap_factor=%.2f, sorted=%s" % (f"{a.docs:,}", a.dim, a.k, a.query_cut, a.heap_factor, bool(a.sorted)))


def shm_dir() -> Path:
    return Path("/dev/shm") if Path("/dev/shm").is_dir() else Path("/tmp")


def prepare_index(tag_items, build_fn, local_rank, world, barrier, keep=False):
    """Local rank 0 builds the index (and, for N > 1, saves it to /dev/shm); the other ranks mmap it."""
    from seismic_b200 import HostIndex
    tag = hashlib.sha1(json.dumps(tag_items).encode()).hexdigest()[:12]
    path = shm_dir() / f"seismic_b200_{tag}.idx"
    timings = {}
    index = None
    if local_rank == 0:
        if not path.exists():
            index = build_fn(timings)
            if world > 1 or keep:
                t = time.time()
                tmp = str(path) + ".tmp%d" % os.getpid()
                index.save(tmp)
                os.replace(tmp, path)
                timings["save_s"] = round(time.time() - t, 2)
        else:
            log(f"reusing {path}")
    barrier()
    if index is None:
        index = HostIndex.load(str(path))
    return index, path, timings


def build_main_index(a, timings):
    from seismic_b200 import Dataset, HostIndex
    cfg = Dataset.synth_config(a.docs, dim=a.dim)
    t = time.time()
    docs = Dataset.synth_documents(cfg)
    timings["gen_s"] = round(time.time() - t, 2)
    log(f"generated {len(docs):,} docs, {docs.nnz:,} nnz in {timings['gen_s']} s")
    t = time.time()
    index = HostIndex.build(docs, n_postings=a.n_postings, centroid_fraction=a.centroid_fraction,
                            summary_energy=a.summary_energy, max_fraction=a.max_fraction)
    timings["build_s"] = round(time.time() - t, 2)
    log(f"built index in {timings['build_s']} s: {index.space_usage()}")
    return index


def build_lv_index(a, timings):
    from seismic_b200 import Dataset, HostIndex
    cfg = Dataset.synth_config(a.lv_docs, dim=a.lv_dim, doc_nnz_mean=a.lv_nnz)
    t = time.time()
    docs = Dataset.synth_documents(cfg)
    timings["gen_s"] = round(time.time() - t, 2)
    log(f"LV: generated {len(docs):,} docs, {docs.nnz:,} nnz in {timings['gen_s']} s")
    t = time.time()
    index = HostIndex.build(docs, n_postings=a.n_postings, centroid_fraction=a.centroid_fraction,
                            summary_energy=a.summary_energy, max_fraction=a.max_fraction, comp_bits=32)
    timings["build_s"] = round(time.time() - t, 2)
    log(f"LV: built index in {timings['build_s']} s: {index.space_usage()}")
    return index


def slice_queries(queries, lo, hi):
    off = queries.offsets
    q_off = (off[lo:hi + 1] - off[lo]).astype(np.uint64)
    return q_off, queries.comps[int(off[lo]):int(off[hi])].copy(), queries.values[int(off[lo]):int(off[hi])].copy()


def run_reference(a, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be built in
    this image (no cargo/rustc; un-vendored git deps), so this is the C++ restatement in oracle/ (kind "port"),
    all host threads, the SAME batch of queries per step as the b200 arm's config 2."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    from seismic_b200 import Dataset
    index, path, timings = prepare_index([a.docs, a.dim, a.n_postings, a.centroid_fraction, a.summary_energy,
                                          a.max_fraction], lambda t: build_main_index(a, t), 0, 1, lambda: None,
                                         keep=a.keep_index)
    queries = Dataset.synth_queries(Dataset.synth_config(a.docs, dim=a.dim), a.queries)
    q_off, q_c, q_v = slice_queries(queries, 0, a.queries)
    n = len(q_off) - 1
    cores = os.cpu_count() or 1
    times = []
    for i in range(a.warmup + a.steps):
        t = time.perf_counter()
        oracle.batch_search(index.view, q_off, q_c, q_v, a.k, a.query_cut, a.heap_factor, first_sorted=bool(a.sorted),
                            n_threads=cores)
        dt = time.perf_counter() - t
        if i >= a.warmup:
            times.append(dt)
    tot = sum(times)
    qps = n * len(times) / tot
    out = {
        "impl": "reference", "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * tot / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "queries_per_gpu_per_step": n, "queries_per_step": n},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"the whole batch of {n} queries per step, all {cores} host threads, C++ oracle "
                                   "(restatement of the Rust path; the Rust crate cannot be compiled here)"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(out)
    if not a.keep_index:
        try:
            os.remove(path)
        except OSError:
            pass


class Runner:
    """Times one (index, query batch, parameters) configuration: device-resident loop with CUDA events and an L2
    flush between steps, then the host-buffer C-ABI loop by wall clock; both max over ranks."""

    def __init__(self, a, rank, local_rank, world, dev, barrier):
        import torch
        self.a, self.rank, self.local_rank, self.world, self.dev, self.barrier = a, rank, local_rank, world, dev, barrier
        self.torch = torch
        self.flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
        self.stream = torch.cuda.current_stream(dev)

    def flush_l2(self, v):
        """Evict L2 between timed steps: write 512 MB.  (Reading another 256 MB afterwards, so that no dirty lines are
        left for the step to write back, was measured: the step's first small kernels got 0.1 ms SLOWER, not faster.)"""
        self.flush.fill_(v)

    def device_buffers(self, q_off, q_c, q_v, k):
        torch = self.torch
        nq = len(q_off) - 1
        return {
            "nq": nq, "k": k,
            "off": torch.from_numpy(q_off.astype(np.int64)).to(self.dev),
            "c": torch.from_numpy(q_c.astype(np.int32)).to(self.dev),
            "v": torch.from_numpy(q_v).to(self.dev),
            "ids": torch.empty((nq, k), dtype=torch.int64, device=self.dev),
            "sc": torch.empty((nq, k), dtype=torch.float32, device=self.dev),
            "cnt": torch.empty(nq, dtype=torch.int32, device=self.dev),
        }

    def search_device(self, gpu, b, cut, hf, srt):
        return gpu.batch_search_device(b["off"].data_ptr(), b["c"].data_ptr(), b["v"].data_ptr(), b["nq"], b["k"], cut,
                                       hf, b["ids"].data_ptr(), b["sc"].data_ptr(), b["cnt"].data_ptr(),
                                       first_sorted=srt)

    def timed(self, step_fn, steps, warmup):
        """W warm-up + K timed steps, each bracketed by CUDA events on the launching stream; flush_l2() between
        steps (outside the brackets) evicts L2.  Returns (sum of event ms as max over ranks, per-step stats)."""
        torch = self.torch
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for _ in range(max(warmup, 0)):
            self.flush_l2(1)
            step_fn()
        self.barrier()
        torch.cuda.synchronize()
        stats = []
        t_wall = time.perf_counter()
        for i in range(steps):
            self.flush_l2(i & 0xFF)
            ev[i][0].record(self.stream)
            stats.append(step_fn())
            ev[i][1].record(self.stream)
        torch.cuda.synchronize()
        self.barrier()
        t_wall = time.perf_counter() - t_wall
        ms = sum(s.elapsed_time(e) for s, e in ev)
        return self.max_over_ranks(ms), stats, t_wall

    def timed_wall(self, step_fn, steps, warmup):
        """End-to-end steps by wall clock (each step ends with the results on the host)."""
        times = []
        for i in range(max(warmup, 1) + steps):
            if i == max(warmup, 1):
                self.barrier()
            t = time.perf_counter()
            step_fn()
            dt = time.perf_counter() - t
            if i >= max(warmup, 1):
                times.append(dt)
        return self.max_over_ranks(1e3 * sum(times)), len(times)

    def max_over_ranks(self, x):
        if self.world == 1:
            return x
        import torch.distributed as dist
        tt = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())


def parity_block(got, ref, nq):
    ids, scores, counts = got
    o_ids, o_sc, o_cnt = ref[:3]
    mism = int(((ids != o_ids).any(axis=1) | (counts != o_cnt)).sum())
    return {"queries": int(nq), "id_mismatch_queries": mism,
            "scores_bit_identical": bool(np.array_equal(scores.view(np.uint32), o_sc.view(np.uint32))),
            "max_abs_score_diff": float(np.nanmax(np.abs(np.where(np.isfinite(o_sc), scores - o_sc, 0.0)))) if nq else 0.0}


def load_traffic(name, **match):
    """Per-launch DRAM bytes of k_search from the committed ncu capture of exactly this workload (or None)."""
    try:
        tr = json.loads((REPO / "profiles" / name).read_text())
        for k, v in match.items():
            tv = tr.get(k)
            if isinstance(v, float):
                if tv is None or abs(tv - v) > 1e-6:
                    return None
            elif tv != v:
                return None
        return tr["dram_bytes_per_launch"]
    except Exception:
        return None


def roofline_block(alg_bytes, ms_search, traffic, peaks, extra=None):
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes / (ms_search * 1e-3) / 1e9
    out = {"bound": "hbm", "kernel": "k_search", "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": achieved / peak, "traffic": traffic,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback, profiling recipe)",
           "algorithmic_bytes_per_launch": int(alg_bytes), "ms_per_launch": ms_search}
    if extra:
        out.update(extra)
    return out


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if a.impl == "reference":
        return run_reference(a, rank, world)

    import torch
    import torch.distributed as dist
    from seismic_b200 import Dataset, GpuIndex, pinned_array, recall_at_k
    from seismic_b200.distributed import gather_results, pack_results, shard_bounds, unpack_results
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the b200 arm has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    want = set(x.strip() for x in a.configs.split(",") if x.strip())
    peaks = {}
    try:
        peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    cores = os.cpu_count() or 1
    oracle = None
    if rank == 0:
        import oracle as _oracle
        _oracle.build()
        oracle = _oracle
    R = Runner(a, rank, local_rank, world, dev, barrier)
    k, nq, srt = a.k, a.queries, bool(a.sorted)

    # ------------------------------------------------------------------ config 2 (top level): weak scaling
    index, path, timings = prepare_index([a.docs, a.dim, a.n_postings, a.centroid_fraction, a.summary_energy,
                                          a.max_fraction], lambda t: build_main_index(a, t), local_rank, world, barrier,
                                         keep=a.keep_index)
    # every rank derives the same query set and takes its own slice (weak scaling: a.queries per GPU)
    queries = Dataset.synth_queries(Dataset.synth_config(a.docs, dim=a.dim), nq * world)
    q_off, q_c, q_v = slice_queries(queries, rank * nq, (rank + 1) * nq)
    t = time.time()
    gpu = GpuIndex(index, local_rank)
    timings["upload_s"] = round(time.time() - t, 2)
    gpu.set_stream(R.stream.cuda_stream)
    log(f"rank {rank}: image {gpu.device_bytes / 1e9:.2f} GB in HBM, upload {timings['upload_s']} s")
    B = R.device_buffers(q_off, q_c, q_v, k)
    gathered = [torch.empty((nq, k * 3 + 1), dtype=torch.int32, device=dev) for _ in range(world)] \
        if world > 1 and rank == 0 else None

    def step_weak(cut=a.query_cut, hf=a.heap_factor):
        st = R.search_device(gpu, B, cut, hf, srt)
        if world > 1:  # single NCCL gather of the result tuples (ids u64 as 2 x i32, score bits, counts)
            dist.gather(pack_results(B["ids"], B["sc"], B["cnt"]), gathered, dst=0)
        return st

    sampler = ClockSampler(local_rank)
    step_weak()  # first call: allocations
    if rank == 0:
        sampler.start()
    dev_ms, stats, t_wall = R.timed(step_weak, a.steps, a.warmup)
    gathered_host = [g.cpu() for g in gathered] if gathered is not None else None
    # e2e: host buffers through the public C-ABI call, H2D + D2H inside, wall clock
    gpu.set_stream(0)
    res_box = {}

    # page-locked host buffers for the queries and the results (the C ABI copies them by DMA, no staging copy)
    P = {"off": pinned_array(q_off.shape, np.uint64), "c": pinned_array(q_c.shape, np.uint32),
         "v": pinned_array(q_v.shape, np.float32),
         "out": (pinned_array((nq, k), np.uint64), pinned_array((nq, k), np.float32), pinned_array(nq, np.uint32))}
    P["off"][:], P["c"][:], P["v"][:] = q_off, q_c, q_v

    def step_e2e(cut=a.query_cut, hf=a.heap_factor):
        res_box["res"] = gpu.batch_search(P["off"], P["c"], P["v"], k, cut, hf, first_sorted=srt, out=P["out"])

    e2e_ms, e2e_n = R.timed_wall(step_e2e, a.steps, a.warmup)
    clocks = sampler.stop() if rank == 0 else None
    res = res_box["res"]
    gpu.set_stream(R.stream.cuda_stream)

    out = None
    ex = None
    nr = min(a.recall_queries, nq)
    if rank == 0:
        # full-size parity + algorithmic bytes: the oracle replays the reference's decisions on the whole batch
        # (for N > 1: on the batches of ALL ranks, against what the NCCL gather delivered)
        all_off, all_c, all_v = slice_queries(queries, 0, nq * world)
        t = time.perf_counter()
        o_all = oracle.batch_search(index.view, all_off, all_c, all_v, k, a.query_cut, a.heap_factor, first_sorted=srt,
                                    n_threads=cores)
        t_all = time.perf_counter() - t
        o_ids, o_sc, o_cnt, _ = o_all
        ref0 = (o_ids[:nq], o_sc[:nq], o_cnt[:nq])
        par = parity_block(res, ref0, nq)
        dev_ids = B["ids"].cpu().numpy().view(np.uint64)
        par["id_mismatch_queries_device_api"] = int((dev_ids != ref0[0]).any(axis=1).sum())
        if gathered_host is not None:
            g_ids, g_sc, g_cnt = unpack_results(torch.cat(gathered_host, dim=0), k)
            gp = parity_block((g_ids.numpy().view(np.uint64), g_sc.numpy(), g_cnt.numpy().view(np.uint32)),
                              (o_ids, o_sc, o_cnt), nq * world)
            par["gathered_all_ranks"] = gp
        # the oracle's byte counters for rank 0's batch alone (what one k_search launch of rank 0 processes)
        ost = oracle.batch_search(index.view, q_off, q_c, q_v, k, a.query_cut, a.heap_factor, first_sorted=srt,
                                  n_threads=cores)[3] if world > 1 else o_all[3]
        # single-thread CPU sample, perf_inverted_index protocol
        n1 = min(a.cpu_sample, nq)
        s_off = q_off[: n1 + 1]
        t = time.perf_counter()
        oracle.batch_search(index.view, s_off, q_c[: int(s_off[-1])], q_v[: int(s_off[-1])], k, a.query_cut,
                            a.heap_factor, first_sorted=srt, n_threads=1)
        t_1 = time.perf_counter() - t
        recall = None
        if nr > 0:
            r_off = q_off[: nr + 1]
            ex = gpu.exact_search(r_off, q_c[: int(r_off[-1])], q_v[: int(r_off[-1])], k)
            recall = recall_at_k(ex[0], ex[2], res[0][:nr], res[2][:nr])
        ms_search = float(np.mean([s["ms_search"] for s in stats]))
        alg_search = ost["bytes_postings"] + ost["bytes_forward"] + ost["bytes_query_out"]
        traffic = load_traffic("r2_traffic.json", docs=a.docs, queries=nq, k=k, query_cut=a.query_cut, sorted=a.sorted,
                               heap_factor=float(a.heap_factor))
        qps = world * nq * a.steps / (dev_ms * 1e-3)
        e2e_qps = world * nq * e2e_n / (e2e_ms * 1e-3)
        alg_step = ost["bytes_total"]
        out = {
            "metric": "queries/sec", "value": qps, "unit": "queries/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": dev_ms / a.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "queries_per_gpu_per_step": nq, "parallelism": f"replicas x{world}",
                       "index": {"n_postings": a.n_postings, "centroid_fraction": a.centroid_fraction,
                                 "summary_energy": a.summary_energy, "max_fraction": a.max_fraction, "values": "f16"},
                       "l2": "512 MB buffer written between timed steps; a step gathers %.1f GB from a %.1f GB image"
                             % (stats[-1]["fwd_bytes"] / 1e9, gpu.device_bytes / 1e9)},
            "e2e": {"value": e2e_qps, "unit": "queries/s",
                    "h2d_bytes_per_step": int(q_off.nbytes + q_c.nbytes + q_v.nbytes),
                    "d2h_bytes_per_step": int(res[0].nbytes + res[1].nbytes + res[2].nbytes)},
            "gpu_launches": int(sum(s["n_launches"] for s in stats)),
            "roofline": roofline_block(alg_search, ms_search, traffic, peaks,
                                       {"read_bytes_per_launch": int(stats[-1]["fwd_bytes"]),
                                        "whole_step": {"algorithmic_bytes": int(alg_step), "ms": dev_ms / a.steps,
                                                       "frac": alg_step / (dev_ms / a.steps * 1e-3) / 1e9 /
                                                       float(peaks.get("hbm_gbs", 6650.0))}}),
            "cpu_baseline": {"value": nq * world / t_all, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"{nq * world} queries once, {cores} host threads, C++ oracle",
                             "single_thread": {"value": n1 / t_1, "unit": "queries/s", "cores": 1,
                                               "us_per_query": 1e6 * t_1 / n1, "sample": f"first {n1} queries"}},
            "clocks": clocks,
            "parity": par,
            "recall_at_k": recall, "recall_queries": nr,
            "kernel_ms": {k2: float(np.mean([s[k2] for s in stats])) for k2 in
                          ("ms_prep", "ms_summary", "ms_search", "ms_finish", "ms_total")},
            "work": {"docs_scored_gpu": int(stats[-1]["docs_scored"]), "docs_scored_reference": int(ost["docs_scored"]),
                     "blocks_scored_gpu": int(stats[-1]["blocks_scored"]),
                     "blocks_evaluated_reference": int(ost["blocks_evaluated"]),
                     "algorithmic_bytes_per_query": ost["bytes_total"] / nq},
            "phase_share": (lambda c: [round(x / max(1, sum(c)), 4) for x in c])(stats[-1]["phase_cycles"]),
            "waves_per_query": round(stats[-1].get("waves", 0) / nq, 2), "ctas_per_sm": stats[-1].get("ctas_per_sm"),
            "setup_s": timings, "wall_s_timed_region": t_wall,
            "configs": {},
        }
    sub_steps = max(5, min(a.steps, 20))

    def sub_config(label, g, view, off, c, v, kk, cut, hf, alg_kind="search", extra_cfg=None):
        """Device-timed + e2e + parity for one more configuration on this rank's batch (weak: every rank runs it)."""
        Bx = R.device_buffers(off, c, v, kk)
        gx = [torch.empty((len(off) - 1, kk * 3 + 1), dtype=torch.int32, device=dev) for _ in range(world)] \
            if world > 1 and rank == 0 else None

        def st_dev():
            s = R.search_device(g, Bx, cut, hf, srt)
            if world > 1:
                dist.gather(pack_results(Bx["ids"], Bx["sc"], Bx["cnt"]), gx, dst=0)
            return s

        st_dev()
        ms, sts, _ = R.timed(st_dev, sub_steps, 3)
        g.set_stream(0)
        box = {}

        n_sub = len(off) - 1
        pin = {"off": pinned_array(off.shape, np.uint64), "c": pinned_array(c.shape, np.uint32),
               "v": pinned_array(v.shape, np.float32),
               "out": (pinned_array((n_sub, kk), np.uint64), pinned_array((n_sub, kk), np.float32),
                       pinned_array(n_sub, np.uint32))}
        pin["off"][:], pin["c"][:], pin["v"][:] = off, c, v

        def st_e2e():
            box["res"] = g.batch_search(pin["off"], pin["c"], pin["v"], kk, cut, hf, first_sorted=srt, out=pin["out"])

        ems, en = R.timed_wall(st_e2e, sub_steps, 2)
        g.set_stream(R.stream.cuda_stream)
        if rank != 0:
            return None, box["res"]
        n = len(off) - 1
        t0 = time.perf_counter()
        ref = oracle.batch_search(view, off, c, v, kk, cut, hf, first_sorted=srt, n_threads=cores)
        t_ref = time.perf_counter() - t0
        ost2 = ref[3]
        ms_s = float(np.mean([s["ms_search"] for s in sts]))
        alg = ost2["bytes_postings"] + ost2["bytes_forward"] + ost2["bytes_query_out"]
        blk = {
            "value": world * n * sub_steps / (ms * 1e-3), "unit": "queries/s", "ms_per_step": ms / sub_steps,
            "steps": sub_steps, "scaling": "weak",
            "e2e": {"value": world * n * en / (ems * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": int(off.nbytes + c.nbytes + v.nbytes),
                    "d2h_bytes_per_step": int(sum(x.nbytes for x in box["res"]))},
            "roofline": roofline_block(alg, ms_s, None, peaks, {"read_bytes_per_launch": int(sts[-1]["fwd_bytes"])}),
            "parity": parity_block(box["res"], ref, n),
            "cpu_baseline": {"value": n / t_ref, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"the batch of {n} queries once, {cores} host threads, C++ oracle"},
            "kernel_ms": {k2: float(np.mean([s[k2] for s in sts])) for k2 in
                          ("ms_prep", "ms_summary", "ms_search", "ms_finish", "ms_total")},
            "work": {"docs_scored_gpu": int(sts[-1]["docs_scored"]), "docs_scored_reference": int(ost2["docs_scored"]),
                     "algorithmic_bytes_per_query": ost2["bytes_total"] / n},
            "phase_share": (lambda cc: [round(x / max(1, sum(cc)), 4) for x in cc])(sts[-1]["phase_cycles"]),
            "gpu_launches": int(sum(s["n_launches"] for s in sts)),
        }
        if extra_cfg:
            blk["config"] = extra_cfg
        return blk, box["res"]

    # ------------------------------------------------------------------ the metric's operating point (recall >= 0.97)
    if "r97" in want and nr > 0:
        cand_list = [x for x in a.r97.split(",") if x]
        pick = torch.zeros(1, dtype=torch.int64, device=dev)
        tried = []
        if rank == 0:
            r_off = q_off[: nr + 1]
            chosen = len(cand_list) - 1
            for i, cand in enumerate(cand_list):
                cut97, hf97 = int(cand.split(":")[0]), float(cand.split(":")[1])
                gpu.set_stream(0)
                r = gpu.batch_search(r_off, q_c[: int(r_off[-1])], q_v[: int(r_off[-1])], k, cut97, hf97, first_sorted=srt)
                gpu.set_stream(R.stream.cuda_stream)
                rec = recall_at_k(ex[0], ex[2], r[0], r[2])
                tried.append({"query_cut": cut97, "heap_factor": hf97, "recall_at_k": rec})
                if rec >= 0.97:
                    chosen = i
                    break
            pick[0] = chosen
        if world > 1:
            dist.broadcast(pick, src=0)
        cand = cand_list[int(pick.item())]
        cut97, hf97 = int(cand.split(":")[0]), float(cand.split(":")[1])
        blk, r97res = sub_config("r97", gpu, index.view, q_off, q_c, q_v, k, cut97, hf97)
        if rank == 0:
            blk["query_cut"], blk["heap_factor"] = cut97, hf97
            blk["recall_at_k"] = recall_at_k(ex[0], ex[2], r97res[0][:nr], r97res[2][:nr])
            blk["recall_queries"] = nr
            blk["candidates_tried"] = tried
            out["configs"]["r97"] = blk
            out["at_recall_0.97"] = {"query_cut": cut97, "heap_factor": hf97, "recall_at_k": blk["recall_at_k"],
                                     "value": blk["value"], "e2e": blk["e2e"]["value"], "unit": "queries/s",
                                     "parity_id_mismatch_queries": blk["parity"]["id_mismatch_queries"]}

    # ------------------------------------------------------------------ config 3: ONE batch split over the GPUs
    if "3" in want:
        # strong scaling: the first `nq` queries (rank 0's weak batch) are the batch; rank r searches its contiguous
        # slice and the tuples are gathered on rank 0 (one NCCL gather); rank 0 checks ALL nq results
        lo, hi = shard_bounds(nq, rank, world)
        s_off, s_c, s_v = slice_queries(queries, lo, hi)
        Bs = R.device_buffers(s_off, s_c, s_v, k)
        gbox = {}

        def step_strong():
            st = R.search_device(gpu, Bs, a.query_cut, a.heap_factor, srt)
            if world > 1:
                gbox["g"] = gather_results(Bs["ids"], Bs["sc"], Bs["cnt"], nq, dst=0)
            else:
                gbox["g"] = (Bs["ids"], Bs["sc"], Bs["cnt"])
            return st

        step_strong()
        s_ms, s_stats, _ = R.timed(step_strong, sub_steps, 3)
        # end to end: pinned host queries -> device, search, gather, gathered tuples -> host (rank 0)
        h_off = torch.from_numpy(s_off.astype(np.int64)).pin_memory()
        h_c = torch.from_numpy(s_c.astype(np.int32)).pin_memory()
        h_v = torch.from_numpy(s_v).pin_memory()
        hres = {}

        def step_strong_e2e():
            Bs["off"].copy_(h_off, non_blocking=True)
            Bs["c"].copy_(h_c, non_blocking=True)
            Bs["v"].copy_(h_v, non_blocking=True)
            step_strong()
            if rank == 0:
                g = gbox["g"]
                hres["r"] = (g[0].cpu(), g[1].cpu(), g[2].cpu())
            torch.cuda.synchronize()

        se_ms, se_n = R.timed_wall(step_strong_e2e, sub_steps, 2)
        if rank == 0:
            g_ids, g_sc, g_cnt = hres["r"]
            got = (g_ids.numpy().view(np.uint64), g_sc.numpy(), g_cnt.numpy().view(np.uint32))
            par3 = parity_block(got, ref0, nq)
            ms_parts = {k2: float(np.mean([s[k2] for s in s_stats])) for k2 in
                        ("ms_prep", "ms_summary", "ms_search", "ms_finish", "ms_total")}
            out["configs"]["3_strong"] = {
                "value": nq * sub_steps / (s_ms * 1e-3), "unit": "queries/s", "ms_per_step": s_ms / sub_steps,
                "steps": sub_steps, "scaling": "strong", "n_gpus": world, "queries_per_step_total": nq,
                "queries_per_gpu": hi - lo,
                "e2e": {"value": nq * se_n / (se_ms * 1e-3), "unit": "queries/s",
                        "h2d_bytes_per_step": int(h_off.numel() * 8 + h_c.numel() * 4 + h_v.numel() * 4),
                        "d2h_bytes_per_step": int(nq * (k * 12 + 4)), "includes": "H2D, search, NCCL gather, D2H on rank 0"},
                "parity": par3,
                "kernel_ms_rank0": ms_parts,
                "fixed_cost_ms": {"rank0_kernels_other_than_search": ms_parts["ms_total"] - ms_parts["ms_search"],
                                  "step_minus_rank0_kernels": s_ms / sub_steps - ms_parts["ms_total"]},
                "gpu_launches": int(sum(s["n_launches"] for s in s_stats)),
            }
        del Bs

    # ------------------------------------------------------------------ config 4: DotVByte forward index
    if "4" in want:
        t = time.time()
        vb = index.convert_to_dotvbyte()
        t_conv = time.time() - t
        del gpu
        torch.cuda.empty_cache()
        gvb = GpuIndex(vb, local_rank)
        gvb.set_stream(R.stream.cuda_stream)
        fwd16 = index.space_usage()["forward"]
        blk, vres = sub_config("4_dotvbyte", gvb, vb.view, q_off, q_c, q_v, k, a.query_cut, a.heap_factor)
        if rank == 0:
            blk["config"] = {"workload": workload_name(a) + ", DotVByte forward index",
                             "forward_bytes": vb.space_usage()["forward"], "forward_bytes_f16": fwd16,
                             "image_bytes": gvb.device_bytes, "convert_s": round(t_conv, 2)}
            if ex is not None:
                blk["recall_at_k_vs_f16_exact"] = recall_at_k(ex[0], ex[2], vres[0][:nr], vres[2][:nr])
            blk["vs_f16_same_run"] = blk["value"] / out["value"]
            out["configs"]["4_dotvbyte"] = blk
        del gvb, vb
    else:
        del gpu
    del index
    torch.cuda.empty_cache()
    barrier()
    if local_rank == 0 and not a.keep_index:
        try:
            os.remove(path)
        except OSError:
            pass

    # ------------------------------------------------------------------ config 5 shape: large vocabulary, k = 100
    if "5" in want:
        lv_timings = {}
        lv, lv_path, lv_timings = prepare_index(["lv", a.lv_docs, a.lv_dim, a.lv_nnz, a.n_postings, a.centroid_fraction,
                                                 a.summary_energy, a.max_fraction],
                                                lambda t: build_lv_index(a, t), local_rank, world, barrier, keep=a.keep_index)
        lq = Dataset.synth_queries(Dataset.synth_config(a.lv_docs, dim=a.lv_dim, doc_nnz_mean=a.lv_nnz), nq * world)
        l_off, l_c, l_v = slice_queries(lq, rank * nq, (rank + 1) * nq)
        t = time.time()
        glv = GpuIndex(lv, local_rank)
        lv_timings["upload_s"] = round(time.time() - t, 2)
        glv.set_stream(R.stream.cuda_stream)
        blk, lres = sub_config("5_lv", glv, lv.view, l_off, l_c, l_v, a.lv_k, a.query_cut, a.heap_factor)
        if rank == 0:
            blk["config"] = {"workload": "synthetic large-vocabulary: %s docs, vocab %d (u32 components), ~150 nnz/doc, "
                                         "k=%d, query_cut=%d, heap_factor=%.2f; BASELINE configs[4] names 50 M docs — "
                                         "the CPU index build limits a bench run to this size (DESIGN.md)"
                                         % (f"{a.lv_docs:,}", a.lv_dim, a.lv_k, a.query_cut, a.heap_factor),
                             "image_bytes": glv.device_bytes, "space": lv.space_usage(), "setup_s": lv_timings}
            n5 = min(200, nq)
            o5 = l_off[: n5 + 1]
            t0 = time.perf_counter()
            ex5 = oracle.exact_search(lv.view, o5, l_c[: int(o5[-1])], l_v[: int(o5[-1])], a.lv_k) if a.lv_docs <= 2_000_000 else None
            if ex5 is not None:
                blk["recall_at_k"] = recall_at_k(ex5[0], ex5[2], lres[0][:n5], lres[2][:n5])
                blk["recall_queries"] = n5
            out["configs"]["5_lv"] = blk
        del glv, lv
        barrier()
        if local_rank == 0 and not a.keep_index:
            try:
                os.remove(lv_path)
            except OSError:
                pass

    if rank == 0:
        emit(out)
    barrier()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
