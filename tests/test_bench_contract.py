"""bench.py contract on the CPU side: the reference arm (`--impl reference`) needs no GPU, prints ONE JSON line with the
keys the driver reads, and never touches /root/reference at run time."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--batch", "128"], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in rec, key
    assert rec["impl"] == "reference" and rec["higher_is_better"] is True and rec["vs_baseline"] is None
    assert rec["config"]["workload"] == "cfg2" and rec["dtype"] == "f64" and rec["data"] == "synthetic"
    assert rec["cpu_baseline"]["kind"] in ("port", "reference") and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["d2h_bytes_per_step"] == 0
    assert rec["e2e"]["value"] == rec["value"] and rec["value"] > 0


def test_nothing_reads_the_reference_tree_at_run_time():
    """/root/reference does not exist on the GPU box: the product package, bench.py and the entry points must not
    mention it (only the oracle shim and the golden generator, which run in the build container, do)."""
    offenders = []
    for base, _, files in os.walk(ROOT):
        if any(part in base for part in (".git", "gpurun_out", "__pycache__", os.path.join("tests", "golden"), "build")):
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".h")):
                continue
            path = os.path.join(base, f)
            rel = os.path.relpath(path, ROOT)
            if rel in (os.path.join("oracle", "ref_shim.py"), os.path.join("tests", "test_bench_contract.py")):
                continue
            if "/root/reference" in open(path, errors="ignore").read():
                offenders.append(rel)
    assert offenders == [], offenders
