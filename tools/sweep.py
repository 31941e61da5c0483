#!/usr/bin/env python
"""Build the benchmark index once and sweep (query_cut, heap_factor): recall@k vs exact, QPS, bytes (GPU box)."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from seismic_b200 import Dataset, GpuIndex, HostIndex, recall_at_k

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=8_800_000)
ap.add_argument("--queries", type=int, default=10000)
ap.add_argument("--recall-queries", type=int, default=1000)
ap.add_argument("--n-postings", type=int, default=3500)
ap.add_argument("--centroid-fraction", type=float, default=0.1)
ap.add_argument("--summary-energy", type=float, default=0.4)
ap.add_argument("--max-fraction", type=float, default=1.5)
ap.add_argument("--k", type=int, default=10)
a = ap.parse_args()
cfg = Dataset.synth_config(a.docs)
t = time.time(); docs = Dataset.synth_documents(cfg); print("gen", round(time.time() - t, 1), flush=True)
t = time.time(); index = HostIndex.build(docs, n_postings=a.n_postings, centroid_fraction=a.centroid_fraction,
                                         summary_energy=a.summary_energy, max_fraction=a.max_fraction)
print("build", round(time.time() - t, 1), index.space_usage(), flush=True)
del docs
q = Dataset.synth_queries(cfg, a.queries)
gpu = GpuIndex(index, 0)
nr = min(a.recall_queries, a.queries)
r_off = q.offsets[: nr + 1]
ex = gpu.exact_search(r_off, q.comps[: int(r_off[-1])], q.values[: int(r_off[-1])], a.k)
print("exact ms", gpu.last_stats, flush=True)
rows = []
for cut in (3, 4, 5, 6, 8, 10):
    for hf in (0.8, 0.9, 1.0):
        for srt in (True,):
            best = None
            for rep in range(3):
                ids, sc, cnt = gpu.batch_search(q.offsets, q.comps, q.values, a.k, cut, hf, first_sorted=srt)
                st = dict(gpu.last_stats)
                if best is None or st["ms_total"] < best["ms_total"]:
                    best = st
            rec = recall_at_k(ex[0], ex[2], ids[:nr], cnt[:nr])
            row = {"query_cut": cut, "heap_factor": hf, "sorted": srt, "recall": round(rec, 4),
                   "qps_kernels": round(a.queries / best["ms_total"] * 1e3), "ms_search": round(best["ms_search"], 3),
                   "ms_total": round(best["ms_total"], 3), "docs_per_query": round(best["docs_scored"] / a.queries, 1)}
            rows.append(row)
            print(json.dumps(row), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/sweep_%d.json" % a.docs).write_text(json.dumps(rows, indent=1))
