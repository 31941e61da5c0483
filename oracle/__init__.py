"""ORACLE — test infrastructure only (see oracle/oracle_search.cpp).  ctypes loader for liboracle.so.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs;
seismic_b200/ never imports this package."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "_build" / "liboracle.so"

ORDER_LANES8, ORDER_SEQ = 0, 1


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "n_queries", "lists_visited", "blocks_total", "blocks_evaluated", "postings_seen", "docs_scored", "results",
        "bytes_summaries", "bytes_postings", "bytes_forward", "bytes_query_out", "bytes_total")] + [("seconds", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


def build(force: bool = False, native: bool = False) -> Path:
    src = HERE / "oracle_search.cpp"
    if force or not LIB.exists() or LIB.stat().st_mtime < src.stat().st_mtime:
        args = ["make", "-C", str(HERE), "-B"]
        if native:
            args.append("CXXFLAGS=-O3 -march=native -ffp-contract=off -fno-fast-math -std=c++17 -fPIC -pthread")
        res = subprocess.run(args, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.oracle_batch_search.restype = C.c_int
        _lib.oracle_batch_search.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.POINTER(OracleStats)]
        _lib.oracle_summary_distances.restype = C.c_int
        _lib.oracle_summary_distances.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        _lib.oracle_exact_search.restype = C.c_int
        _lib.oracle_exact_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _batch(offsets, comps, values):
    # layout of SgpuQueryBatch {u64 n, ptr, ptr, ptr}
    class QB(C.Structure):
        _fields_ = [("n", C.c_uint64), ("o", C.c_void_p), ("c", C.c_void_p), ("v", C.c_void_p)]
    return QB(len(offsets) - 1, offsets.ctypes.data, comps.ctypes.data, values.ctypes.data)


def _params(k, query_cut, heap_factor, n_knn, first_sorted):
    class SP(C.Structure):
        _fields_ = [("k", C.c_uint32), ("cut", C.c_uint32), ("hf", C.c_float), ("nknn", C.c_uint32), ("fs", C.c_int32)]
    return SP(k, query_cut, heap_factor, n_knn, 1 if first_sorted else 0)


def batch_search(view, offsets, comps, values, k, query_cut, heap_factor, n_knn=0, first_sorted=True,
                 order=ORDER_LANES8, n_threads=1, n_runs=1):
    """`view` is a ctypes SgpuIndexView (seismic_b200._native.IndexView). Returns ids, scores, counts, stats."""
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    comps = np.ascontiguousarray(comps, dtype=np.uint32)
    values = np.ascontiguousarray(values, dtype=np.float32)
    nq = len(offsets) - 1
    ids = np.empty((nq, k), dtype=np.uint64)
    scores = np.empty((nq, k), dtype=np.float32)
    counts = np.empty(nq, dtype=np.uint32)
    qb = _batch(offsets, comps, values)
    p = _params(k, query_cut, heap_factor, n_knn, first_sorted)
    st = OracleStats()
    rc = lib().oracle_batch_search(C.addressof(view), C.addressof(qb), C.addressof(p), order, n_threads, n_runs,
                                   ids.ctypes.data, scores.ctypes.data, counts.ctypes.data, C.byref(st))
    if rc == -1:
        raise ValueError("oracle: invalid argument (k == 0, unsorted query or component >= dim)")
    if rc != 0:
        raise NotImplementedError("oracle: unsupported configuration (rc=%d)" % rc)
    return ids, scores, counts, st.as_dict()


def summary_distances(view, list_id, comps, values, n_blocks):
    comps = np.ascontiguousarray(comps, dtype=np.uint32)
    values = np.ascontiguousarray(values, dtype=np.float32)
    out = np.zeros(n_blocks, dtype=np.float32)
    rc = lib().oracle_summary_distances(C.addressof(view), list_id, comps.ctypes.data, values.ctypes.data, len(comps),
                                        out.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_summary_distances rc=%d" % rc)
    return out


def exact_search(view, offsets, comps, values, k, n_threads=0):
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    comps = np.ascontiguousarray(comps, dtype=np.uint32)
    values = np.ascontiguousarray(values, dtype=np.float32)
    nq = len(offsets) - 1
    ids = np.empty((nq, k), dtype=np.uint64)
    scores = np.empty((nq, k), dtype=np.float32)
    counts = np.empty(nq, dtype=np.uint32)
    qb = _batch(offsets, comps, values)
    rc = lib().oracle_exact_search(C.addressof(view), C.addressof(qb), k, n_threads, ids.ctypes.data,
                                   scores.ctypes.data, counts.ctypes.data)
    if rc != 0:
        raise ValueError("oracle_exact_search rc=%d" % rc)
    return ids, scores, counts
