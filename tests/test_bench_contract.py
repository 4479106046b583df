"""bench.py contract pieces that run without a GPU: the reference arm (CPU port on the host cores) prints one JSON line
with the required keys; the algorithmic-FLOP bookkeeping matches SURVEY.md section 8(d)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["higher_is_better"] is True
    assert line["metric"].startswith("DPA loss fwd+bwd tokens/s")
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == dict(value=line["value"], unit="tokens/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_algorithmic_flops_match_survey():
    from halva_b200.synth import CONFIGS, algorithmic_flops
    cfg2 = CONFIGS["cfg2"]
    assert abs(algorithmic_flops(cfg2, True) - 85.9e12) / 85.9e12 < 2e-3          # (98304*3 + 32768) * 262.144 MF
    assert abs(algorithmic_flops(cfg2, False) - 60.1e12) / 60.1e12 < 2e-3
    assert abs(algorithmic_flops(CONFIGS["cfg1"], True) - 2.68e12) / 2.68e12 < 3e-3
    assert cfg2.tokens == 98304
