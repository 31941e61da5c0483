#!/usr/bin/env python
"""kNN graph at scale (GPU box): time build_knn (N self-searches as GPU batches), then recall / time of searches
with and without Knn::refine, and parity of the refine against the CPU oracle on a sample."""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from seismic_b200 import Dataset, HostIndex, api, recall_at_k

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1_000_000)
ap.add_argument("--queries", type=int, default=10000)
ap.add_argument("--nknn", type=int, default=10)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--recall-queries", type=int, default=1000)
ap.add_argument("--parity-queries", type=int, default=1000)
a = ap.parse_args()

cfg = Dataset.synth_config(a.docs)
docs = Dataset.synth_documents(cfg)
host = HostIndex.build(docs); del docs
q = Dataset.synth_queries(cfg, a.queries)
idx = api.SeismicIndexRaw(host)
gpu = idx.gpu
t = time.time(); idx.build_knn(a.nknn); t_build = time.time() - t
out = {"docs": a.docs, "nknn": a.nknn, "build_knn_s": round(t_build, 2),
       "self_searches_per_s": round(a.docs / t_build), "graph_MB": round(host.knn.nbytes / 1e6, 1),
       "full_rows": float((host.knn != np.uint64(0xFFFFFFFFFFFFFFFF)).all(axis=1).mean())}
print(json.dumps(out), flush=True)
nr = min(a.recall_queries, a.queries)
r_off = q.offsets[: nr + 1]
ex = gpu.exact_search(r_off, q.comps[: int(r_off[-1])], q.values[: int(r_off[-1])], a.k)
rows = []
for cut, hf in ((1, 0.9), (2, 0.9), (3, 0.8), (5, 0.9)):
    for n_knn in (0, a.nknn):
        best = None
        for _ in range(3):
            ids, sc, cnt = gpu.batch_search(q.offsets, q.comps, q.values, a.k, cut, hf, n_knn=n_knn, first_sorted=True)
            st = dict(gpu.last_stats)
            if best is None or st["ms_total"] < best["ms_total"]:
                best = st
        row = {"query_cut": cut, "heap_factor": hf, "n_knn": n_knn, "recall": round(recall_at_k(ex[0], ex[2], ids[:nr], cnt[:nr]), 4),
               "ms_total": round(best["ms_total"], 3), "ms_search": round(best["ms_search"], 3),
               "docs_per_query": round(best["docs_scored"] / a.queries, 1)}
        rows.append(row)
        print(json.dumps(row), flush=True)
import oracle
n = min(a.parity_queries, a.queries)
o = q.offsets[: n + 1]
ref = oracle.batch_search(host.view, o, q.comps[: int(o[-1])], q.values[: int(o[-1])], a.k, 3, 0.8, n_knn=a.nknn, first_sorted=True)
ids, sc, cnt = gpu.batch_search(o, q.comps[: int(o[-1])], q.values[: int(o[-1])], a.k, 3, 0.8, n_knn=a.nknn, first_sorted=True)
out["parity"] = {"queries": n, "id_mismatch_queries": int(((ids != ref[0]).any(axis=1) | (cnt != ref[2])).sum()),
                 "scores_bit_identical": bool(np.array_equal(sc.view(np.uint32), ref[1].view(np.uint32)))}
out["rows"] = rows
print(json.dumps(out["parity"]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/knn_bench_%d.json" % a.docs).write_text(json.dumps(out, indent=1))
