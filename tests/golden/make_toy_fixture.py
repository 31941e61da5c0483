"""Generates tests/golden/toy_dataset.json from the reference's examples/toy_dataset (20 SPLADE docs, 5 queries).

Run in the build container only (it reads /root/reference); the GPU box only sees the committed fixture:
    python tests/golden/make_toy_fixture.py
Token ids follow the reference's .bin conversion rule (rank in the sorted token set,
scripts/convert_json_to_inner_format.py:188-190) so the fixture is reproducible.  The fixture also stores the
ORACLE's results for the BASELINE configs[0] search (k=10, query_cut=3, heap_factor=0.8, sorted=True): the
reference ships no golden search output for this data set, so these are regression vectors for the oracle and
parity vectors for the GPU path, not reference outputs."""
import json
import sys
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(REPO))
SRC = Path("/root/reference/examples/toy_dataset")


def main():
    from seismic_b200 import Dataset, HostIndex
    import oracle
    docs = [json.loads(l) for l in (SRC / "documents.jsonl").read_text().splitlines() if l.strip()]
    queries = [json.loads(l) for l in (SRC / "queries.jsonl").read_text().splitlines() if l.strip()]
    tokens = sorted({t for d in docs for t in d["vector"]})
    tid = {t: i for i, t in enumerate(tokens)}

    def enc(rec, drop_unknown):
        items = sorted((tid[t], float(np.float32(v))) for t, v in rec["vector"].items() if not drop_unknown or t in tid)
        return [i for i, _ in items], [v for _, v in items]
    d_enc = [enc(d, False) for d in docs]
    q_enc = [enc(q, True) for q in queries]
    ds = Dataset.from_lists([np.array(c, np.uint32) for c, _ in d_enc], [np.array(v, np.float32) for _, v in d_enc],
                            dim=len(tokens))
    index = HostIndex.build(ds)
    off = np.zeros(len(q_enc) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c, _ in q_enc])
    qc = np.concatenate([np.array(c, np.uint32) for c, _ in q_enc])
    qv = np.concatenate([np.array(v, np.float32) for _, v in q_enc])
    out = {"dim": len(tokens), "doc_ids": [str(d["id"]) for d in docs], "query_ids": [str(q["id"]) for q in queries],
           "docs": d_enc, "queries": q_enc, "results": {}}
    for name, (k, cut, hf, srt) in {"k10_cut3_hf0.8_sorted": (10, 3, 0.8, True), "k10_cut3_hf0.8_unsorted": (10, 3, 0.8, False),
                                     "k5_cut10_hf1.0_sorted": (5, 10, 1.0, True)}.items():
        ids, scores, counts, _ = oracle.batch_search(index.view, off, qc, qv, k, cut, hf, first_sorted=srt)
        out["results"][name] = {"k": k, "query_cut": cut, "heap_factor": hf, "sorted": srt,
                                "ids": [ids[i, :counts[i]].tolist() for i in range(len(counts))],
                                "scores": [[float(x) for x in scores[i, :counts[i]]] for i in range(len(counts))]}
    ex = oracle.exact_search(index.view, off, qc, qv, 10)
    out["exact_top10"] = [ex[0][i, :ex[2][i]].tolist() for i in range(len(ex[2]))]
    (Path(__file__).parent / "toy_dataset.json").write_text(json.dumps(out))
    print("wrote", Path(__file__).parent / "toy_dataset.json", "docs", len(docs), "queries", len(queries), "dim", len(tokens))


if __name__ == "__main__":
    main()
