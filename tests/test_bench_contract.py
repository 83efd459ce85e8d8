"""CPU checks of bench.py's contract pieces that need no GPU: the algorithmic FLOP counts the roofline uses, the
reference arm's JSON line (run here on a tiny configuration), and the arguments the driver passes."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

# algorithmic FLOP / sample (2 MAC, MLPs only)
SURVEY_FLOPS = {"cfg1_toy": 2113536, "cfg2_power": 5365760, "cfg3_miniboone": 23633920, "cfg4_hepmass": 89128960,
                "cfg5_bsds300": 183336960}


@pytest.mark.parametrize("name", list(SURVEY_FLOPS))
def test_flops_per_sample_match_survey(name):
    assert bench.flops_per_sample(bench.CONFIGS[name]) == SURVEY_FLOPS[name]


def test_peaks_have_the_keys_the_roofline_uses():
    pk = bench.peaks()
    assert set(pk) >= {"bf16_burst", "bf16_sustained", "hbm", "source"}
    assert pk["bf16_burst"] >= pk["bf16_sustained"] > 0 and pk["hbm"] > 0


def test_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` is the CPU arm (oracle port on the host threads): one JSON line with the base contract's keys,
    impl = reference, a cpu_baseline describing the run and an e2e object that repeats the line's own value."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg1_toy",
                          "--steps", "1", "--warmup", "1", "--batch", "256"], capture_output=True, text=True, env=env,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["gpu_launches"] == 0
    staged = os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "models", "boosted_flow.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_port_fallback():
    """Without the staged reference (or with --port) the arm times the oracle's torch-CPU port and says so."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--port", "--config", "cfg1_toy",
                          "--steps", "1", "--warmup", "1", "--batch", "256"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_staged_reference_matches_its_source():
    if not os.path.isdir("/root/reference"):
        pytest.skip("the reference tree only exists in the build container")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "vendor_reference.py"), "--check"], capture_output=True,
                         text=True)
    if not os.path.isdir(os.path.join(ROOT, "baseline", "_ref")):
        pytest.skip("baseline/_ref not staged yet (python __graft_entry__.py build stages it)")
    assert out.returncode == 0, out.stdout + out.stderr


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""
