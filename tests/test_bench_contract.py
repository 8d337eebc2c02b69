"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the driver's contract."""
import json
import os
import subprocess
import sys

import helpers


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(helpers.REPO, "bench.py"), "--impl", "reference", "--workload", "furnace", "--steps", "2", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "Mpaths/s" and line["unit"] == "Mpaths/s"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True and line["value"] > 0
    assert line["config"]["workload"].startswith("configs[0]")
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    p = subprocess.run([sys.executable, os.path.join(helpers.REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "furnace", "--steps", "1"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_b200_arm_refuses_to_run_without_a_device():
    import torch

    if torch.cuda.is_available():
        return
    p = subprocess.run([sys.executable, os.path.join(helpers.REPO, "bench.py"), "--workload", "furnace", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
