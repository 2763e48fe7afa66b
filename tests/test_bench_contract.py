"""bench.py contract pieces that do not need a GPU: the reference arm prints exactly one
JSON line with the required keys, and the synthetic-workload helpers are deterministic."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_reference_arm_prints_one_json_line_with_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "MP/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"] == bench.METRIC and "workload" in d["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_synthetic_workload_helpers():
    wm = bench.make_wm_map()
    assert wm.shape == (135, 240) and wm.dtype == np.uint8 and set(np.unique(wm)) == {0, 255}
    assert (wm[:, :58] == 255).all() and (wm[:6] == 255).all()          # white padding around the QR-like core
    assert np.array_equal(wm, bench.make_wm_map())
    kinds = [bench.image_kind(i) for i in range(8)]
    assert kinds.count("natural") == 4 and kinds.count("random") == 2 and kinds.count("regions") == 2
    a = bench.cpu_image(5, 64)
    assert a.shape == (64, 1920, 3) and a.dtype == np.uint8 and np.array_equal(a, bench.cpu_image(5, 64))
    assert abs(bench.ALGO_BYTES_PER_PX - 6.015625) < 1e-12


def test_clock_sampler_parses_nvidia_smi_rows():
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    s.lines = [(10.0, "0, 1965, 1965, 600.1, 0x0000000000000004, Not Active, Not Active, Not Active, Active"),
               (10.1, "0, 1800, 1965, 900.5, 0x0000000000000004, Not Active, Not Active, Not Active, Active"),
               (10.2, "0, 1950, 1965, 950.0, 0x0000000000000000, Not Active, Not Active, Not Active, Not Active")]
    out = s.stop(9.9, 10.3)
    assert out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0 and out["reasons"] == ["sw_power_cap"]


@pytest.mark.gpu
def test_b200_arm_prints_one_json_line_with_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "3", "--warmup", "3",
                        "--images", "32", "--e2e-images", "8", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "MP/s" and d["value"] > 1000 and d["n_gpus"] == 1 and d["gpu_launches"] == 3
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-3
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 8 * 1080 * 1920 * 3 and e["d2h_bytes_per_step"] == 8 * 1080 * 1920 * 3
    assert e["matches_device_path"] is True and d["extract"]["watermark_bits_recovered_on_natural_images"] is True
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
