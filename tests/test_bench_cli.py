"""bench.py plumbing that can run without a GPU: the reference arm end to end on a tiny corpus, and a static check that
every `self.<attr>` the Runner reads is assigned in its __init__ (the GPU arm itself only runs on the B200 box)."""
import ast
import json
import os
import subprocess
import sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]


def test_reference_arm_runs_on_cpu():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--docs", "20000", "--dim", "4000",
                          "--queries", "200", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["gpu_launches"] == 0
    assert d["config"]["queries_per_step"] == 200


def test_runner_attributes_are_initialised():
    tree = ast.parse((REPO / "bench.py").read_text())
    runner = next(n for n in ast.walk(tree) if isinstance(n, ast.ClassDef) and n.name == "Runner")
    init = next(n for n in runner.body if isinstance(n, ast.FunctionDef) and n.name == "__init__")
    stored = {n.attr for n in ast.walk(init) if isinstance(n, ast.Attribute) and isinstance(n.ctx, ast.Store)
              and isinstance(n.value, ast.Name) and n.value.id == "self"}
    methods = {n.name for n in runner.body if isinstance(n, ast.FunctionDef)}
    later = set()
    used = {n.attr for n in ast.walk(tree) if isinstance(n, ast.Attribute) and isinstance(n.ctx, ast.Load)
            and isinstance(n.value, ast.Name) and n.value.id in ("self", "R")}
    # attributes read through `self.` / `R.` anywhere in bench.py must exist on the Runner (other classes use disjoint names)
    other = {"device", "rows", "proc", "t", "Q", "_read"}  # ClockSampler
    missing = {u for u in used if u not in stored | methods | later | other}
    assert not missing, missing
