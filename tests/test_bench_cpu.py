"""
CPU: the parts of bench.py that need no GPU -- the work counts behind `roofline`, the north-star check, the workload
table, and the reference arm end to end (its JSON line is what the driver parses for the speed-up ratio).
"""

import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def test_work_counts_of_the_likelihood_sweep():
    # A_mufu = 980 O + 3276, A_flop = 8232 O + 63220 per unit
    assert bench.algorithmic_work(3)["survey_8d"] == {"mufu_ops": 6216, "fp32_flop": 87916}
    assert bench.algorithmic_work(64)["survey_8d"] == {"mufu_ops": 65996, "fp32_flop": 590068}
    # executed counts of the single-bin row form (profiles/r2_sass_ksmogn_o1.txt: 34 FFMA2 + 16 FADD2 + 7 FMUL2 + 2 FADD
    # per pixel pair, 98 pairs, 600 lane operations of prologue / epilogue per patch)
    w = bench.algorithmic_work(1)
    assert w["fp32_lane_ops"] == 98 * (2 * (34 + 16 + 7) + 2) + 600 == 11968
    assert w["fp32_flop"] == 98 * (4 * 34 + 2 * (16 + 7) + 2) + 900 == 18932
    assert w["mufu_ops"] == 196 + 3 * 196 * 2 + 2 * 2 * 14 == 1428
    # more bins never cost less
    assert bench.algorithmic_work(3)["mufu_ops"] > w["mufu_ops"]
    assert bench.algorithmic_work(64)["mufu_ops"] > bench.algorithmic_work(4)["mufu_ops"]


def test_north_star_check_is_throughput_over_the_survey_roofline():
    peaks = {"fma": 35.0e12, "mufu": 4.5e12}   # lane-ops/s (an FMA = 2 flops), MUFU ops/s
    roof3 = 1.0 / max(6216 / 4.5e12, 87916 / 70.0e12)
    roof1 = 1.0 / max(4256 / 4.5e12, 71452 / 70.0e12)
    out = bench.north_star_check(peaks, 9.0e8, {"c3_o3": {"value": 4.5e8}})
    assert out["survey_8d_step_roofline"]["o3_aoi_frames_per_s"] == pytest.approx(roof3)
    assert out["survey_8d_step_roofline"]["o1_aoi_frames_per_s"] == pytest.approx(roof1)
    assert out["c3_merged_bins_over_o3_roofline"] == pytest.approx(9.0e8 / roof3)
    assert out["c3_three_bins_kept_over_o3_roofline"] == pytest.approx(4.5e8 / roof3)
    assert "c3_three_bins_kept_over_o3_roofline" not in bench.north_star_check(peaks, 9.0e8, None)


def test_workload_table_matches_the_baseline_configs():
    w = bench.WORKLOADS
    assert w["c1"][:4] == (5, 100, 5, 100)            # BASELINE.json configs[0]
    assert w["c2"][:4] == (100, 1000, 100, 1000)      # configs[1]
    assert w["c2mb"][:4] == (100, 1000, 10, 512)      # the reference's default minibatch (main.py:1428-1431)
    assert w["c3"][:4] == (1000, 5000, 1000, 5000)    # configs[2], the headline
    assert w["c4"][:2] == (500, 2000) and bench.WORKLOAD_CHANNELS["c4"] == 2        # configs[3]
    assert w["c5"][:2] == (200, 2000) and bench.WORKLOAD_MODEL["c5"] == "cosmos+hmm"   # configs[4]
    for n, shard in ((2, "c3s2"), (4, "c3s4"), (8, "c3s8")):   # one rank's share of the headline at N GPUs
        assert w[shard][0] * n == w["c3"][0] and w[shard][1] == w["c3"][1]
    cfg = bench.config_of("c3", 8, "strong", 0, False)
    assert cfg["aois_per_gpu"] == 125 and cfg["aois_total"] == 1000 and "1000 AOIs" in cfg["workload"]


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference`: the oracle on the host cores, one JSON line with the native arm's metric / config."""
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "AOI-frames/s" and line["higher_is_better"] is True
    assert line["metric"] == "cosmos SVI AOI-frames/sec (ELBO fwd+bwd+Adam)"
    assert line["config"] == bench.config_of("c3", 1, "strong", 0, False)
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "AOI-frames/s", "h2d_bytes_per_step": 0,
                                                  "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "10 AOIs x 512 frames" in cb["sample"]
