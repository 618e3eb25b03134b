"""bench.py's stdout contract, on the arm that runs without a GPU: `--impl reference` (the reference's CPU path,
oracle/_ref when it was compiled here, else the oracle port) prints exactly ONE line, a JSON object with the keys
the driver reads; ranks other than 0 print nothing and exit 0."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CMD = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
       "--reads-per-sample", "2000", "--samples-per-gpu", "2"]


def test_reference_arm_prints_one_json_line():
    res = subprocess.run(CMD, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, res.stdout[:2000]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "aligned_bases_per_sec" and line["unit"] == "aligned bases/s"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["value"] > 0 and line["ms_per_step"] > 0
    assert line["config"]["workload"].startswith("cfg2_summarise")
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "reference+port", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run(CMD, capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
