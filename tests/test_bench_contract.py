"""bench.py --impl reference (the CPU leg the driver runs beside our arm) must print exactly one JSON line with the
contract's keys.  Runs the oracle port once at full size on this container's cores (about half a minute)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "images/sec ViT-L 896x448 MIM train step"
    assert j["unit"] == "images/s" and j["higher_is_better"] is True and j["vs_baseline"] is None
    assert j["value"] > 0 and abs(j["value"] * j["ms_per_step"] / 1e3 - 1.0) < 1e-6      # B = 1 per step
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["unit"] == j["unit"]
    assert e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0
    assert "workload" in j["config"] and j["n_gpus"] == 1 and j["steps"] == 1


def test_reference_arm_is_silent_on_non_zero_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
