"""bench.py contract pieces that need no GPU: the reference arm (CPU port of the reference's path, bounded sample) prints ONE
JSON line with the keys the driver reads; the measured arm's helpers do not import oracle/."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "denoiser_sample_steps_per_sec" and d["unit"] == "sample-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "configs[1]" in d["config"]["workload"] and "model" not in d["config"]


def test_measured_arm_does_not_import_the_oracle():
    """only cpu_reference_run (the cpu_baseline / --impl reference leg) may touch oracle/"""
    src = open(os.path.join(ROOT, "bench.py")).read()
    body = src.split("def cpu_reference_run", 1)
    before = body[0]
    after = body[1].split("\ndef main", 1)[1]
    assert not re.search(r"^\s*(from|import)\s+oracle", before, flags=re.M)
    assert not re.search(r"^\s*(from|import)\s+oracle", after, flags=re.M)
    for f in os.listdir(os.path.join(ROOT, "lidarcrafter_b200")):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(ROOT, "lidarcrafter_b200", f)).read().replace("C oracle", ""), f
