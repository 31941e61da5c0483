#!/usr/bin/env python
"""bench.py — queries/sec of the Seismic query hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one pass of the hot path (SeismicIndex.batch_search, Python default sorted=True) over one batch of
`--queries` synthetic queries per GPU against the synthetic SPLADE-v3-shaped corpus of BASELINE.json configs[1]
(8.8 M docs, vocab 30 522, ~120 nnz/doc, ~40 nnz/query, k=10, query_cut=3, heap_factor=0.8; index built with the
reference's Python defaults).  The index is replicated on every GPU and each rank owns its own batch (weak
scaling); results are gathered on rank 0 with one NCCL gather per step.

  value  = whole-job queries/s, inputs resident in HBM, timed with CUDA events on the launching stream
  e2e    = the same through the host-buffer C-ABI call (H2D of the queries + D2H of the results inside)
  roofline / cpu_baseline / clocks as described in DESIGN.md §Measurement.
Rank 0 prints ONE JSON line.
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
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--docs", type=int, default=int(os.environ.get("SEISMIC_BENCH_DOCS", 8_800_000)))
    ap.add_argument("--dim", type=int, default=30522)
    ap.add_argument("--queries", type=int, default=int(os.environ.get("SEISMIC_BENCH_QUERIES", 10_000)),
                    help="queries per GPU per step")
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--query-cut", type=int, default=3)
    ap.add_argument("--heap-factor", type=float, default=0.8)
    ap.add_argument("--sorted", type=int, default=1)
    ap.add_argument("--n-postings", type=int, default=3500)
    ap.add_argument("--centroid-fraction", type=float, default=0.1)
    ap.add_argument("--summary-energy", type=float, default=0.4)
    ap.add_argument("--max-fraction", type=float, default=1.5)
    ap.add_argument("--recall-queries", type=int, default=1000, help="queries used for recall@k vs exact (0 = skip)")
    ap.add_argument("--cpu-sample", type=int, default=2000, help="queries of the CPU baseline sample")
    ap.add_argument("--wave-docs", type=int, default=0)
    ap.add_argument("--first-wave-docs", type=int, default=0)
    ap.add_argument("--r97", default="5:0.9,6:0.9,5:0.8,8:0.9",
                    help="query_cut:heap_factor candidates, cheapest first, for the extra run that must reach "
                         "recall@10 >= 0.97 (the first that does is reported); empty = skip")
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


def prepare(a, rank, local_rank, world, barrier):
    """Rank 0 of the node generates the corpus, builds the index and saves it to /dev/shm; the others mmap it."""
    from seismic_b200 import Dataset, HostIndex
    tag = hashlib.sha1(json.dumps([a.docs, a.dim, a.n_postings, a.centroid_fraction, a.summary_energy,
                                   a.max_fraction]).encode()).hexdigest()[:12]
    shm = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path("/tmp")
    path = shm / f"seismic_b200_{tag}.idx"
    cfg = Dataset.synth_config(a.docs, dim=a.dim)
    timings = {}
    index = None
    if local_rank == 0:
        if not path.exists():
            t = time.time()
            docs = Dataset.synth_documents(cfg)
            timings["gen_s"] = round(time.time() - t, 2)
            log(f"generated {len(docs):,} docs, {docs.nnz:,} nnz in {timings['gen_s']} s")
            t = time.time()
            index = HostIndex.build(docs, n_postings=a.n_postings, centroid_fraction=a.centroid_fraction,
                                    summary_energy=a.summary_energy, max_fraction=a.max_fraction)
            timings["build_s"] = round(time.time() - t, 2)
            log(f"built index in {timings['build_s']} s: {index.space_usage()}")
            del docs
            if world > 1 or a.keep_index:
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
    # every rank derives the same query set and takes its own slice (weak scaling: a.queries per GPU)
    queries = Dataset.synth_queries(cfg, a.queries * world)
    off = queries.offsets
    lo, hi = rank * a.queries, (rank + 1) * a.queries
    q_off = (off[lo:hi + 1] - off[lo]).astype(np.uint64)
    q_c = queries.comps[int(off[lo]):int(off[hi])].copy()
    q_v = queries.values[int(off[lo]):int(off[hi])].copy()
    return index, (q_off, q_c, q_v), path, timings


def run_reference(a, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The Rust crate cannot be built in
    this image (no cargo/rustc; un-vendored git deps), so this is the C++ restatement in oracle/ (kind "port"),
    all host threads, one bounded sample of the same workload per step."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    index, (q_off, q_c, q_v), path, timings = prepare(a, 0, 0, 1, lambda: None)
    n = min(a.cpu_sample, len(q_off) - 1)
    s_off = q_off[: n + 1]
    s_c, s_v = q_c[: int(s_off[-1])], q_v[: int(s_off[-1])]
    cores = os.cpu_count() or 1
    times = []
    for i in range(a.warmup + a.steps):
        t = time.perf_counter()
        oracle.batch_search(index.view, s_off, s_c, s_v, a.k, a.query_cut, a.heap_factor, first_sorted=bool(a.sorted),
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
        "config": {"workload": workload_name(a), "queries_per_step": n},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": f"first {n} queries of the batch per step, all {cores} host threads, C++ oracle "
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


def main():
    a = parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if a.impl == "reference":
        return run_reference(a, rank, world)

    import torch
    import torch.distributed as dist
    from seismic_b200 import GpuIndex, recall_at_k
    from seismic_b200.distributed import pack_results
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

    index, (q_off, q_c, q_v), path, timings = prepare(a, rank, local_rank, world, barrier)
    nq, k = len(q_off) - 1, a.k
    t = time.time()
    gpu = GpuIndex(index, local_rank)
    timings["upload_s"] = round(time.time() - t, 2)
    if a.wave_docs:
        gpu.set_option("wave_docs", a.wave_docs)
    if a.first_wave_docs:
        gpu.set_option("first_wave_docs", a.first_wave_docs)
    stream = torch.cuda.current_stream(dev)
    gpu.set_stream(stream.cuda_stream)
    log(f"rank {rank}: image {gpu.device_bytes / 1e9:.2f} GB in HBM, upload {timings['upload_s']} s")

    # device-resident inputs / outputs
    d_off = torch.from_numpy(q_off.astype(np.int64)).to(dev)
    d_c = torch.from_numpy(q_c.astype(np.int32)).to(dev)
    d_v = torch.from_numpy(q_v).to(dev)
    d_ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
    d_sc = torch.empty((nq, k), dtype=torch.float32, device=dev)
    d_cnt = torch.empty(nq, dtype=torch.int32, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    gathered = None
    if world > 1 and rank == 0:
        gathered = [torch.empty((nq, k * 3 + 1), dtype=torch.int32, device=dev) for _ in range(world)]

    def step_device():
        st = gpu.batch_search_device(d_off.data_ptr(), d_c.data_ptr(), d_v.data_ptr(), nq, k, a.query_cut,
                                     a.heap_factor, d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(),
                                     first_sorted=bool(a.sorted))
        if world > 1:  # single NCCL gather of the result tuples (ids u64 as 2 x i32, scores bits, counts)
            dist.gather(pack_results(d_ids, d_sc, d_cnt), gathered if rank == 0 else None, dst=0)
        return st

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    for _ in range(max(a.warmup, 0)):
        flush.fill_(1)
        step_device()
    barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    stats = []
    t_wall = time.perf_counter()
    for i in range(a.steps):
        flush.fill_(i & 0xFF)  # evict L2 between timed iterations (outside the event bracket)
        ev[i][0].record(stream)
        stats.append(step_device())
        ev[i][1].record(stream)
    torch.cuda.synchronize()
    barrier()
    t_wall = time.perf_counter() - t_wall
    dev_ms = sum(s.elapsed_time(e) for s, e in ev)
    # e2e: host buffers through the public C-ABI call, H2D + D2H inside, wall clock
    gpu.set_stream(0)
    e2e_times = []
    res = None
    for i in range(max(a.warmup, 1) + a.steps):
        t = time.perf_counter()
        res = gpu.batch_search(q_off, q_c, q_v, k, a.query_cut, a.heap_factor, first_sorted=bool(a.sorted))
        dt = time.perf_counter() - t
        if i >= max(a.warmup, 1):
            e2e_times.append(dt)
    clocks = sampler.stop() if rank == 0 else None
    e2e_ms = 1e3 * sum(e2e_times)

    # max over ranks
    if world > 1:
        tt = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = tt.tolist()

    if rank == 0:
        import oracle
        oracle.build()
        cores = os.cpu_count() or 1
        ids, scores, counts = res
        # full-size parity + algorithmic bytes: the oracle replays the reference's decisions on the whole batch
        t = time.perf_counter()
        o_ids, o_sc, o_cnt, ost = oracle.batch_search(index.view, q_off, q_c, q_v, k, a.query_cut, a.heap_factor,
                                                     first_sorted=bool(a.sorted), n_threads=cores)
        t_all = time.perf_counter() - t
        mism = int(((ids != o_ids).any(axis=1) | (counts != o_cnt)).sum())
        score_ok = bool(np.array_equal(scores, o_sc))
        dev_ids = d_ids.cpu().numpy().view(np.uint64)
        mism_dev = int((dev_ids != o_ids).any(axis=1).sum())
        # single-thread CPU sample, perf_inverted_index protocol
        n1 = min(a.cpu_sample, nq)
        s_off = q_off[: n1 + 1]
        t = time.perf_counter()
        oracle.batch_search(index.view, s_off, q_c[: int(s_off[-1])], q_v[: int(s_off[-1])], k, a.query_cut,
                            a.heap_factor, first_sorted=bool(a.sorted), n_threads=1)
        t_1 = time.perf_counter() - t
        # recall@k vs exact on a subset
        recall = None
        if a.recall_queries > 0:
            nr = min(a.recall_queries, nq)
            r_off = q_off[: nr + 1]
            ex = gpu.exact_search(r_off, q_c[: int(r_off[-1])], q_v[: int(r_off[-1])], k)
            recall = recall_at_k(ex[0], ex[2], ids[:nr], counts[:nr])
        # BASELINE's metric is "queries/sec at recall@10 >= 0.97": the named config (query_cut=3, heap_factor=0.8) is
        # not a tuned point of the reference and reaches ~0.95 on this corpus, so the nearest config
        # that does reach 0.97 is measured too (same batch, device-resident inputs, CUDA events of the library).
        r97 = None
        for cand in filter(None, a.r97.split(",") if a.recall_queries > 0 else []):
            cut97, hf97 = int(cand.split(":")[0]), float(cand.split(":")[1])
            ms = []
            for i in range(3 + min(a.steps, 20)):
                st97 = gpu.batch_search_device(d_off.data_ptr(), d_c.data_ptr(), d_v.data_ptr(), nq, k, cut97, hf97,
                                               d_ids.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr(),
                                               first_sorted=bool(a.sorted))
                if i >= 3:
                    ms.append(st97["ms_total"])
            torch.cuda.synchronize()
            ids97 = d_ids.cpu().numpy().view(np.uint64)
            cnt97 = d_cnt.cpu().numpy().view(np.uint32)
            r97 = {"query_cut": cut97, "heap_factor": hf97, "recall_at_k": recall_at_k(ex[0], ex[2], ids97[:nr], cnt97[:nr]),
                   "recall_queries": nr, "value": world * nq / (float(np.mean(ms)) * 1e-3),
                   "unit": "queries/s (sum of kernel times, per-GPU x n_gpus)", "ms_per_step": float(np.mean(ms)),
                   "ms_search": float(st97["ms_search"])}
            if r97["recall_at_k"] >= 0.97:
                break
        ms_search = float(np.mean([s["ms_search"] for s in stats]))
        ms_kernels = float(np.mean([s["ms_total"] for s in stats]))
        alg_search = ost["bytes_postings"] + ost["bytes_forward"] + ost["bytes_query_out"]
        peaks = {}
        try:
            peaks = json.loads((REPO / "MEASURED_PEAKS.json").read_text())
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = alg_search / (ms_search * 1e-3) / 1e9
        traffic = None  # per-launch DRAM bytes of k_search from the committed ncu capture of this exact workload
        try:
            tr = json.loads((REPO / "profiles" / "r1_traffic.json").read_text())
            if (tr["docs"], tr["queries"], tr["k"], tr["query_cut"], tr["sorted"]) == (a.docs, nq, k, a.query_cut, a.sorted) \
                    and abs(tr["heap_factor"] - a.heap_factor) < 1e-6:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        qps = world * nq * a.steps / (dev_ms * 1e-3)
        e2e_qps = world * nq * len(e2e_times) / (e2e_ms * 1e-3)
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
                    "d2h_bytes_per_step": int(ids.nbytes + scores.nbytes + counts.nbytes)},
            "gpu_launches": int(sum(s["n_launches"] for s in stats)),
            "roofline": {"bound": "hbm", "kernel": "k_search", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (profiling recipe)",
                         "algorithmic_bytes_per_launch": int(alg_search), "ms_per_launch": ms_search,
                         "read_bytes_per_launch": int(stats[-1]["fwd_bytes"])},
            "cpu_baseline": {"value": nq / t_all, "unit": "queries/s", "cores": cores, "kind": "port",
                             "sample": f"whole batch of {nq} queries once, {cores} host threads, C++ oracle",
                             "single_thread": {"value": n1 / t_1, "unit": "queries/s", "cores": 1,
                                               "us_per_query": 1e6 * t_1 / n1, "sample": f"first {n1} queries"}},
            "clocks": clocks,
            "parity": {"queries": nq, "id_mismatch_queries_host_api": mism, "id_mismatch_queries_device_api": mism_dev,
                       "scores_bit_identical": score_ok},
            "recall_at_k": recall, "recall_queries": min(a.recall_queries, nq),
            "at_recall_0.97": r97,
            "kernel_ms": {k2: float(np.mean([s[k2] for s in stats])) for k2 in
                          ("ms_prep", "ms_summary", "ms_search", "ms_finish", "ms_total")},
            "work": {"docs_scored_gpu": int(stats[-1]["docs_scored"]), "docs_scored_reference": int(ost["docs_scored"]),
                     "blocks_scored_gpu": int(stats[-1]["blocks_scored"]), "blocks_evaluated_reference": int(ost["blocks_evaluated"]),
                     "algorithmic_bytes_per_query": ost["bytes_total"] / nq},
            "phase_share": (lambda c: [round(x / max(1, sum(c)), 4) for x in c])(stats[-1]["phase_cycles"]),
            "waves_per_query": round(stats[-1].get("waves", 0) / nq, 2), "ctas_per_sm": stats[-1].get("ctas_per_sm"),
            "setup_s": timings, "wall_s_timed_region": t_wall, "ms_kernels_per_step": ms_kernels,
        }
        emit(out)
    barrier()
    if world > 1:
        dist.destroy_process_group()
    if local_rank == 0 and not a.keep_index:
        try:
            os.remove(path)
        except OSError:
            pass


if __name__ == "__main__":
    main()
