"""ctypes bindings of libseismic_b200.so (the C ABI declared in include/seismic_b200.h) + in-tree build.

The library is compiled in-tree (seismic_b200/_lib/) by `build_native()` with
`nvcc -gencode arch=compute_100a,code=sm_100a`; there is no CPU fallback for the search path: if the
library is missing the import fails loudly, and the sgpu_* calls fail with SGPU_ECUDA without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
REPO = ROOT.parent
LIB_DIR = ROOT / "_lib"
LIB_PATH = LIB_DIR / "libseismic_b200.so"      # the product: CUDA kernels + C ABI (sgpu_*) + the host side (shost_*)
HOST_LIB_PATH = LIB_DIR / "libshost_b200.so"   # the host side alone (dataset, CPU index build, I/O, generator): no CUDA
HOST_SRC = sorted((ROOT / "csrc" / "host").glob("*.cpp"))
CUDA_SRC = sorted((ROOT / "csrc" / "cuda").glob("*.cu"))  # sgpu_api.cu + one translation unit per group of k_search instantiations
HEADERS = (
    sorted((ROOT / "csrc" / "host").glob("*.hpp"))
    + sorted((ROOT / "csrc" / "cuda").glob("*.cuh"))
    + [REPO / "include" / "seismic_b200.h"]
)
OBJ_DIR = LIB_DIR / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-pthread,-O3,-march=x86-64-v3",
]


def _stale() -> bool:
    if not LIB_PATH.exists() or not HOST_LIB_PATH.exists():
        return True
    t = min(LIB_PATH.stat().st_mtime, HOST_LIB_PATH.stat().st_mtime)
    return any(p.stat().st_mtime > t for p in HOST_SRC + CUDA_SRC + HEADERS)


def build_native(force: bool = False, verbose: bool = False) -> Path:
    """Compile the CUDA + host sources into seismic_b200/_lib/libseismic_b200.so for sm_100a.
    Every source file is its own nvcc -c job (run in parallel), then one link."""
    if not force and not _stale():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    LIB_DIR.mkdir(exist_ok=True)
    OBJ_DIR.mkdir(exist_ok=True)
    newest_header = max(p.stat().st_mtime for p in HEADERS)
    extra = os.environ.get("SEISMIC_B200_NVCC_EXTRA", "").split()  # e.g. -DSGPU_...=1 for A/B builds
    jobs = []
    for src in CUDA_SRC + HOST_SRC:
        obj = OBJ_DIR / (src.stem + ".o")
        if force or extra or not obj.exists() or obj.stat().st_mtime < max(src.stat().st_mtime, newest_header):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", "-o", str(obj), str(src)]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on %s:\n%s%s" % (src.name, res.stdout, res.stderr))

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4) or 1) as ex:
        list(ex.map(compile_one, jobs))
    tmp = LIB_DIR / (".build_%d.so" % os.getpid())
    objs = [str(OBJ_DIR / (src.stem + ".o")) for src in CUDA_SRC + HOST_SRC]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC,-pthread", "-o", str(tmp), *objs]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, LIB_PATH)
    # the host-only library: what the CPU legs (reference arm of bench.py, index construction) load — they never map
    # the CUDA code
    tmp = LIB_DIR / (".build_host_%d.so" % os.getpid())
    hobjs = [str(OBJ_DIR / (src.stem + ".o")) for src in HOST_SRC]
    cmd = [os.environ.get("CXX", "g++"), "-shared", "-fPIC", "-pthread", "-o", str(tmp), *hobjs]
    if verbose:
        print(" ".join(cmd), flush=True)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    os.replace(tmp, HOST_LIB_PATH)
    return LIB_PATH


# ---------------------------------------------------------------------------------- structs
class IndexView(C.Structure):
    _fields_ = [
        ("comp_bits", C.c_uint32), ("value_kind", C.c_uint32), ("n_docs", C.c_uint64), ("dim", C.c_uint64),
        ("value_scale", C.c_float), ("knn_dim", C.c_uint32),
        ("fwd_offsets", C.c_void_p), ("fwd_comps", C.c_void_p), ("fwd_values", C.c_void_p), ("fwd_nnz", C.c_void_p),
        ("list_post_start", C.c_void_p), ("postings", C.c_void_p), ("list_blk_start", C.c_void_p),
        ("blk_post_off", C.c_void_p), ("blk_min", C.c_void_p), ("blk_quant", C.c_void_p),
        ("list_sc_start", C.c_void_p), ("sc_comp", C.c_void_p), ("list_ent_start", C.c_void_p),
        ("sc_run_off", C.c_void_p), ("ent_blk", C.c_void_p), ("ent_code", C.c_void_p),
        ("knn_neighbours", C.c_void_p),
    ]


class QueryBatch(C.Structure):
    _fields_ = [("n_queries", C.c_uint64), ("offsets", C.c_void_p), ("comps", C.c_void_p), ("values", C.c_void_p)]


class SearchParams(C.Structure):
    _fields_ = [("k", C.c_uint32), ("query_cut", C.c_uint32), ("heap_factor", C.c_float), ("n_knn", C.c_uint32),
                ("first_sorted", C.c_int32)]


class SearchStats(C.Structure):
    _fields_ = [("ms_total", C.c_float), ("ms_prep", C.c_float), ("ms_summary", C.c_float), ("ms_search", C.c_float),
                ("ms_finish", C.c_float), ("n_launches", C.c_uint32), ("ctas_per_sm", C.c_uint32),
                ("docs_scored", C.c_uint64), ("blocks_scored", C.c_uint64), ("blocks_pushed", C.c_uint64),
                ("fwd_bytes", C.c_uint64), ("phase_cycles", C.c_uint64 * 6), ("waves", C.c_uint64),
                ("select_passes", C.c_uint64)]

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "phase_cycles"}
        d["phase_cycles"] = list(self.phase_cycles)
        return d


class BuildConfig(C.Structure):
    _fields_ = [("pruning", C.c_uint32), ("n_postings", C.c_uint32), ("max_fraction", C.c_float),
                ("blocking", C.c_uint32), ("centroid_fraction", C.c_float), ("min_cluster_size", C.c_uint32),
                ("doc_cut", C.c_uint32), ("block_size", C.c_uint32), ("summarization", C.c_uint32),
                ("summary_energy", C.c_float), ("n_components", C.c_uint32), ("comp_bits", C.c_uint32),
                ("value_kind", C.c_uint32), ("n_threads", C.c_uint32), ("kmeans_seed", C.c_uint64)]


class SynthConfig(C.Structure):
    _fields_ = [("n_docs", C.c_uint64), ("dim", C.c_uint64), ("seed", C.c_uint64), ("n_topics", C.c_uint32),
                ("topic_terms", C.c_uint32), ("doc_nnz_mean", C.c_float), ("doc_nnz_sigma", C.c_float),
                ("query_nnz_mean", C.c_float), ("query_nnz_sigma", C.c_float), ("n_threads", C.c_uint32),
                ("reserved", C.c_uint32)]


VAL_F16, VAL_BF16, VAL_F32, VAL_FIXEDU8, VAL_FIXEDU16, VAL_DOTVBYTE = range(6)
PAD_ID = np.uint64(0xFFFFFFFFFFFFFFFF)

# every symbol include/seismic_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = {
    "sgpu_index_create": (C.c_int, [C.POINTER(IndexView), C.c_int, C.POINTER(C.c_void_p)]),
    "sgpu_index_destroy": (None, [C.c_void_p]),
    "sgpu_index_device_bytes": (C.c_uint64, [C.c_void_p]),
    "sgpu_batch_search": (C.c_int, [C.c_void_p, C.POINTER(QueryBatch), C.POINTER(SearchParams), C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.POINTER(SearchStats)]),
    "sgpu_batch_search_device": (C.c_int, [C.c_void_p, C.POINTER(QueryBatch), C.POINTER(SearchParams), C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.POINTER(SearchStats)]),
    "sgpu_index_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sgpu_index_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "sgpu_index_set_knn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "sgpu_exact_search": (C.c_int, [C.c_void_p, C.POINTER(QueryBatch), C.c_uint32, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.POINTER(C.c_float)]),
    "sgpu_group_create": (C.c_int, [C.POINTER(IndexView), C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "sgpu_group_destroy": (None, [C.c_void_p]),
    "sgpu_group_size": (C.c_int, [C.c_void_p]),
    "sgpu_group_set_knn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "sgpu_group_batch_search": (C.c_int, [C.c_void_p, C.POINTER(QueryBatch), C.POINTER(SearchParams), C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.POINTER(SearchStats), C.POINTER(C.c_float)]),
    "sgpu_host_alloc": (C.c_int, [C.c_uint64, C.POINTER(C.c_void_p)]),
    "sgpu_host_free": (None, [C.c_void_p]),
    "sgpu_last_error": (C.c_char_p, []),
    "sgpu_version": (C.c_char_p, []),
    "shost_default_config": (None, [C.POINTER(BuildConfig)]),
    "shost_default_synth": (None, [C.POINTER(SynthConfig)]),
    "shost_dataset_create": (C.c_int, [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_void_p)]),
    "shost_dataset_read_bin": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "shost_dataset_write_bin": (C.c_int, [C.c_void_p, C.c_char_p]),
    "shost_dataset_destroy": (None, [C.c_void_p]),
    "shost_dataset_len": (C.c_uint64, [C.c_void_p]),
    "shost_dataset_dim": (C.c_uint64, [C.c_void_p]),
    "shost_dataset_nnz": (C.c_uint64, [C.c_void_p]),
    "shost_dataset_offsets": (C.c_void_p, [C.c_void_p]),
    "shost_dataset_comps": (C.c_void_p, [C.c_void_p]),
    "shost_dataset_values": (C.c_void_p, [C.c_void_p]),
    "shost_synth_documents": (C.c_int, [C.POINTER(SynthConfig), C.POINTER(C.c_void_p)]),
    "shost_synth_queries": (C.c_int, [C.POINTER(SynthConfig), C.c_uint64, C.POINTER(C.c_void_p)]),
    "shost_index_build": (C.c_int, [C.c_void_p, C.POINTER(BuildConfig), C.POINTER(C.c_void_p)]),
    "shost_index_convert_dotvbyte": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "shost_index_save": (C.c_int, [C.c_void_p, C.c_char_p]),
    "shost_index_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "shost_index_destroy": (None, [C.c_void_p]),
    "shost_index_view": (C.c_int, [C.c_void_p, C.POINTER(IndexView)]),
    "shost_index_nnz": (C.c_uint64, [C.c_void_p]),
    "shost_index_space_usage": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "shost_index_get_doc": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32,
                                      C.POINTER(C.c_uint32)]),
}

_lib = None
_hlib = None


def hlib() -> C.CDLL:
    """The host-only library (shost_* + sgpu_last_error): datasets, CPU index build, index files, generator."""
    global _hlib
    if _hlib is not None:
        return _hlib
    if os.environ.get("SEISMIC_B200_LIB"):  # A/B runs with one foreign build: take everything from it
        _hlib = lib()
        return _hlib
    if _stale():
        if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            build_native()
        elif not HOST_LIB_PATH.exists():
            raise ImportError("seismic_b200: native library %s is missing and nvcc is not available" % HOST_LIB_PATH)
    handle = C.CDLL(str(HOST_LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        if name.startswith("shost_") or name == "sgpu_last_error":
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
    _hlib = handle
    return handle


def lib() -> C.CDLL:
    """Load (building first if the sources are newer) the native library. Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    override = os.environ.get("SEISMIC_B200_LIB")  # A/B runs of two builds on one GPU box (tools/variants.py)
    if override:
        path = Path(override)
    elif _stale():
        path = LIB_PATH
        if os.path.exists(os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")):
            build_native()
        elif not LIB_PATH.exists():
            raise ImportError(
                "seismic_b200: native library %s is missing and nvcc is not available; "
                "run `python -c 'import __graft_entry__ as g; g.build()'`" % LIB_PATH)
    else:
        path = LIB_PATH
    handle = C.CDLL(str(path))
    for name, (res, args) in SYMBOLS.items():
        if override and not hasattr(handle, name):
            continue  # an older build in an A/B run
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


class SeismicError(RuntimeError):
    pass


_ERR = {-1: ValueError, -2: SeismicError, -3: MemoryError, -4: OSError, -5: NotImplementedError}


def check(rc: int, host: bool = False) -> None:
    """Raise for a non-zero status; `host`: the call went to the host-only library (its own error string)."""
    if rc != 0:
        msg = (hlib() if host else lib()).sgpu_last_error().decode("utf-8", "replace")
        raise _ERR.get(rc, SeismicError)(msg or ("seismic_b200 error %d" % rc))


def hcheck(rc: int) -> None:
    check(rc, host=True)


def ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def np_view(address: int, count: int, dtype) -> np.ndarray:
    """Borrowed numpy view of `count` items at a native address (the owner must stay alive)."""
    dt = np.dtype(dtype)
    if count == 0 or not address:
        return np.empty(0, dtype=dt)
    buf = (C.c_uint8 * (count * dt.itemsize)).from_address(address)
    return np.frombuffer(buf, dtype=dt, count=count)
