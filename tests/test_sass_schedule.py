"""Guard on the machine code of the hot loop of k_search (no GPU needed: cuobjdump reads the built library).

The scoring loop runs at the 64-register limit of 4 CTAs x 256 threads per SM.  One register too many and ptxas sinks
the second document's gathers below the first document's table lookups, which halves the bytes in flight per warp
(measured on B200: 5.3 -> 6.2 ms per 10 k queries).  This test pins the good schedule: the two 256-bit gathers of an
iteration (one per document of the group; four 128-bit ones in builds with SGPU_LD256 = 0) are issued before the first
query-table lookup, and the kernel neither spills nor exceeds 64 registers."""
import re
import shutil
import subprocess

import pytest

from seismic_b200 import _native

KERNEL = "k_searchILi256ELi4ELi2ENS_9ByteQueryENS_7RegHeapENS_5Rec16E"


def _cuobjdump(*args):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        return subprocess.run([exe, *args, str(_native.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("cuobjdump not available")


def test_hot_loop_issues_all_gathers_before_the_first_lookup():
    _native.lib()
    sass = _cuobjdump("-sass")
    start = sass.find("Function : _ZN4sgpu8" + KERNEL)
    assert start >= 0, "benchmark instantiation of k_search not found in the library"
    end = sass.find("Function : ", start + 10)
    ops = [m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", sass[start:end], re.M)]
    is_gather = lambda o: re.match(r"LDG\.E(\.NA|\.ENL2)?\.(128|256)\.CONSTANT", o) is not None  # noqa: E731  (ld.global.nc .v4 / .v8)
    first = next(i for i, o in enumerate(ops) if is_gather(o))
    bits = 0
    for o in ops[first:]:
        if is_gather(o):
            bits += 256 if ".256." in o else 128
        elif o.startswith("LDS.U8"):
            break
    assert bits == 512, "ptxas no longer issues the gathers of a scoring iteration (2 x 32 bytes) back to back (%d bits)" % bits


def test_benchmark_kernel_fits_four_ctas_per_sm_without_spills():
    _native.lib()
    res = _cuobjdump("-res-usage")
    m = re.search(KERNEL + r"[^\n]*\n\s*REG:(\d+) STACK:(\d+)", res)
    assert m, "resource usage of the benchmark kernel not found"
    assert int(m.group(1)) <= 64 and int(m.group(2)) == 0, m.group(0)
