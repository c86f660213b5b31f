"""bench.py's reference arm runs on CPU: check the JSON contract of the line it prints (one line on stdout, the keys
the driver reads).  The GPU arm prints the same keys plus roofline / clocks / gpu_launches (checked on the GPU box by
the round-end run itself)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--entities", "3000", "--edges", "50000", "--relations", "8"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1                                   # exactly one JSON line on stdout
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "edges/s" and j["higher_is_better"] is True
    for key in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["value"] > 0 and j["vs_baseline"] is None and j["data"] == "synthetic" and "workload" in j["config"]
    cb, e2e = j["cpu_baseline"], j["e2e"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    assert e2e["value"] == j["value"] and e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
