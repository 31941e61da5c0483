#!/usr/bin/env python
"""Build the benchmark index once and time k_search under several kernel-option sets (GPU box).

Every option set must return the same ids / score bits as the first one (results never depend on tuning knobs).
usage: tools/variants.py --docs 1000000 "hq_occ=4" "hq_occ=3,hq_wave_docs=640" ...
"""
import argparse, json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from seismic_b200 import Dataset, GpuIndex, HostIndex

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=1_000_000)
ap.add_argument("--queries", type=int, default=10000)
ap.add_argument("--k", type=int, default=10)
ap.add_argument("--cut", type=int, default=3)
ap.add_argument("--hf", type=float, default=0.8)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--out", default="gpurun_out/variants.json")
ap.add_argument("--dim", type=int, default=30522)
ap.add_argument("--comp-bits", type=int, default=16)
ap.add_argument("--doc-nnz-mean", type=float, default=115.0)
ap.add_argument("--dotvbyte", action="store_true")
ap.add_argument("--check-oracle", type=int, default=0, help="compare the first N queries of the first option set with the CPU oracle")
ap.add_argument("opts", nargs="+")
a = ap.parse_args()

cfg = Dataset.synth_config(a.docs, dim=a.dim, doc_nnz_mean=a.doc_nnz_mean)
t = time.time(); docs = Dataset.synth_documents(cfg)
index = HostIndex.build(docs, comp_bits=a.comp_bits); del docs
if a.dotvbyte:
    index = index.convert_to_dotvbyte()
q = Dataset.synth_queries(cfg, a.queries)
print("setup s", round(time.time() - t, 1), flush=True)
gpu = GpuIndex(index, 0)
defaults = {"hq_carveout_pct": 0, "hq_cand_cap": 256, "hq_wave_docs": 768, "hq_first_wave_docs": 128, "bucket": 1, "tma": 0, "wide_heap": 1, "occ32": 4}
base = None
rows = []
for opt in a.opts:
    kv = dict(defaults)
    for item in filter(None, opt.split(",")):
        n, _, v = item.partition("=")
        kv[n.strip()] = int(v)
    try:
        for n, v in kv.items():
            gpu.set_option(n, v)
        best = None
        for _ in range(a.reps):
            ids, sc, cnt = gpu.batch_search(q.offsets, q.comps, q.values, a.k, a.cut, a.hf, first_sorted=True)
            st = dict(gpu.last_stats)
            if best is None or st["ms_search"] < best["ms_search"]:
                best = st
        if base is None:
            base = (ids.copy(), sc.copy(), cnt.copy())
            if a.check_oracle:
                import oracle
                n = min(a.check_oracle, a.queries)
                o = q.offsets[: n + 1]
                ref = oracle.batch_search(index.view, o, q.comps[: int(o[-1])], q.values[: int(o[-1])], a.k, a.cut, a.hf,
                                          first_sorted=True, n_threads=0)
                print(json.dumps({"oracle_mismatch_first_%d" % n: int(((ids[:n] != ref[0]).any(axis=1) | (cnt[:n] != ref[2])).sum()),
                                  "scores_equal": bool(np.array_equal(sc[:n], ref[1]))}), flush=True)
        same = bool(np.array_equal(ids, base[0]) and np.array_equal(sc.view(np.uint32), base[1].view(np.uint32))
                    and np.array_equal(cnt, base[2]))
        tot = float(sum(best["phase_cycles"])) or 1.0
        row = {"opts": opt, "ms_search": round(best["ms_search"], 3), "ms_total": round(best["ms_total"], 3), "ms_summary": round(best["ms_summary"], 3), "ms_prep": round(best["ms_prep"], 3),
               "same_results": same, "docs_scored": best["docs_scored"], "ctas_per_sm": best.get("ctas_per_sm"), "waves": best.get("waves"), "passes": best.get("select_passes"),
               "phase_share": [round(c / tot, 3) for c in best["phase_cycles"]]}
    except Exception as e:  # e.g. a configuration that does not fit shared memory
        row = {"opts": opt, "error": str(e)[:200]}
    rows.append(row)
    print(json.dumps(row), flush=True)
Path(a.out).parent.mkdir(exist_ok=True)
Path(a.out).write_text(json.dumps(rows, indent=1))
