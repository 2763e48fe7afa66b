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
    from oracle import live_reference
    assert d["cpu_baseline"]["kind"] == ("reference" if live_reference.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"] == bench.METRIC and "workload" in d["config"]
    # both arms print the same `config` object (the driver compares them)
    import argparse
    assert d["config"] == bench.make_config(argparse.Namespace(images=1024), 1)


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_synthetic_workload_helpers():
    wm = bench.make_wm_map()
    assert wm.shape == (135, 240) and wm.dtype == np.uint8 and wm.min() == 0 and wm.max() == 255
    assert (wm[:, :52] == 255).all() and (wm[:, -52:] == 255).all()     # white padding left and right of the centred QR
    assert 0.2 < (wm < 128).mean() < 0.35                               # a QR: about half of its modules are dark
    assert np.array_equal(wm, bench.make_wm_map())
    import qr_util as Q
    payload, _ = bench.payload_and_png()
    assert Q.decode_map(wm) == payload                                  # the map itself carries the payload
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
                        "--images", "32", "--no-cpu-baseline", "--sustain-seconds", "0.2"],
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
    assert e["value"] > 0 and e["h2d_bytes_per_step"] >= 32 * 1080 * 1920 * 3 and e["d2h_bytes_per_step"] == 32 * 1080 * 1920 * 3
    assert e["images_per_step_per_gpu"] == 32                      # the e2e leg runs on the same shard as `value`
    assert e["matches_device_path"] is True and d["extract"]["watermark_bits_recovered_on_natural_images"] is True
    assert e["extract"]["value"] > 0 and e["extract"]["h2d_bytes_per_step"] == 2 * 32 * 1080 * 1920 * 3
    assert e["extract"]["matches_device_path"] is True
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert rf["sustained"]["frac"] > 0 and rf["faithful"]["embed_frac"] > 0 and rf["extract"]["frac"] > 0
    c = d["configs"]
    assert set(c) >= {"c1_512_embed_extract", "c2_4k_latency", "c4_extract_with_helper_data", "c5_svd_sweep", "block_sizes"}
    assert all("error" not in (v if isinstance(v, dict) else {}) for v in c.values()), c
    assert c["c1_512_embed_extract"]["payload_decoded_gpu"] is True and c["c2_4k_latency"]["payload_byte_exact"] is True
    nat = c["c4_extract_with_helper_data"]["payload_byte_exact"]["natural"].split("/")
    assert nat[0] == nat[1] and int(nat[1]) == 16
    assert c["c4_extract_with_helper_data"]["helper_data"]["decrypted_text_equals_input"] is True
