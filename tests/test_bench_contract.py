"""CPU: bench.py's reference arm (the reference -- oracle/_ref when present, else the oracle port -- timed on the host cores) prints one JSON line that
carries the contract keys.  Uses the real 1024x2048 workload with a single timed frame."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("frames/sec at 1024x2048") and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["value"] > 0 and d["steps"] == 1 and d["vs_baseline"] is None and d["scaling"] == "weak"
    cb = d["cpu_baseline"]
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "Testing", "model", "pspnet", "td4_psp18.py"))
    assert cb["kind"] == ("reference" if have_ref else "port")      # the reference's own package when oracle/_ref exists
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "1024x2048" in cb["sample"]
    assert d["warmup"] == 3 and d["config"]["config_id"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "td4-psp18 1024x2048" in d["config"]["workload"]


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
