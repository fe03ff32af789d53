"""CPU: bench.py's reference arm prints exactly one JSON line on stdout with the contract's keys
(the GPU arm needs a device; its line is checked on the box by the driver)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--cells", "20000",
                          "--k", "15", "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    rec = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in rec, key
    assert rec["impl"] == "reference" and rec["unit"] == "edges/s" and rec["value"] > 0
    assert rec["cpu_baseline"]["kind"] in ("reference", "port") and rec["cpu_baseline"]["cores"] >= 1
    assert rec["e2e"]["h2d_bytes_per_step"] == 0 and rec["e2e"]["value"] == rec["value"]
    assert "workload" in rec["config"] and "sample" in rec["config"]


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
