"""The C-ABI library loads and exports every symbol include/seismic_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (REPO / "include" / "seismic_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:sgpu|shost)_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(native):
    lib = ctypes.CDLL(str(native.LIB_PATH))
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_bindings_cover_header(native):
    assert set(declared_symbols()) == set(native.SYMBOLS)


def test_version_and_defaults(native):
    lib = native.lib()
    assert b"sm_100a" in lib.sgpu_version()
    cfg = native.BuildConfig()
    lib.shost_default_config(ctypes.byref(cfg))
    # Python defaults of the reference, src/pylib/mod.rs:329
    assert (cfg.n_postings, cfg.min_cluster_size, cfg.doc_cut) == (3500, 2, 15)
    assert abs(cfg.centroid_fraction - 0.1) < 1e-7 and abs(cfg.summary_energy - 0.4) < 1e-7
    assert abs(cfg.max_fraction - 1.5) < 1e-7
    assert ctypes.sizeof(native.BuildConfig) == 64


def test_struct_layouts_match_header(native):
    # sizes the C compiler gives these structs (x86-64 SysV): guards against binding drift
    assert ctypes.sizeof(native.IndexView) == 32 + 17 * 8
    assert ctypes.sizeof(native.QueryBatch) == 32
    assert ctypes.sizeof(native.SearchParams) == 20
    assert ctypes.sizeof(native.SearchStats) == 28 + 4 + 32 + 48 + 16
    assert ctypes.sizeof(native.SynthConfig) == 56


def test_no_gpu_fails_loudly(native, synth_small):
    """Without a CUDA device the product path must raise, never fall back to a CPU implementation."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from seismic_b200 import GpuIndex
    _, _, index = synth_small
    with pytest.raises(native.SeismicError):
        GpuIndex(index, 0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under seismic_b200/ may import, include, link or load it
    (comments citing it as the definition of the summation order are fine)."""
    import re
    for p in (REPO / "seismic_b200").rglob("*"):
        if p.suffix == ".py":
            text = p.read_text()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), p
            assert "liboracle" not in text and "oracle/_build" not in text, p
        elif p.suffix in {".cpp", ".hpp", ".cu", ".cuh", ".h"}:
            text = p.read_text()
            assert not re.search(r"#\s*include[^\n]*oracle", text), p
            assert "liboracle" not in text, p
            if "dlopen" in text:  # the only run-time loaded library is NCCL (multi-GPU gather, sgpu_group_*)
                assert p.name == "sgpu_api.cu" and re.findall(r'"(lib[^"]*\.so[^"]*)"', text) == ["libnccl.so.2", "libnccl.so"], p


def test_group_entry_fails_loudly_without_gpu(native, synth_small):
    """sgpu_group_create: bad arguments are rejected; without a CUDA device it raises (no CPU fallback)."""
    import torch
    from seismic_b200 import GpuGroup
    _, _, index = synth_small
    with pytest.raises(ValueError):
        GpuGroup(index, [])
    with pytest.raises(ValueError):
        GpuGroup(index, [0, 0])
    if not torch.cuda.is_available():
        with pytest.raises(native.SeismicError):
            GpuGroup(index, [0])


def test_cpu_legs_do_not_map_the_product_library():
    """Datasets, the CPU index build and index files live in libshost_b200.so: a process that only uses them (the
    reference arm of bench.py, the oracle's callers) never maps the CUDA library."""
    import subprocess
    import sys
    code = (
        "import sys; sys.path.insert(0, %r)\n"
        "from seismic_b200 import Dataset, HostIndex\n"
        "cfg = Dataset.synth_config(2000, dim=500)\n"
        "idx = HostIndex.build(Dataset.synth_documents(cfg), n_postings=100)\n"
        "q = Dataset.synth_queries(cfg, 4)\n"
        "import oracle; oracle.batch_search(idx.view, q.offsets, q.comps, q.values, 5, 3, 0.8)\n"
        "maps = open('/proc/self/maps').read()\n"
        "print('libseismic_b200.so' in maps, 'libshost_b200.so' in maps)\n" % str(REPO))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout.split()
    assert out == ["False", "True"], out
