"""bench.py's contract where it can be checked without a GPU: the reference arm prints exactly one JSON line with the
required keys (the reference's own progress prints must not leak onto stdout), and the product arm refuses to run
without a CUDA device instead of falling back to a CPU path."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "suzanne640"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--workload", "suzanne640"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--workload", "suzanne640"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CPU fallback" in (out.stderr + out.stdout)


def test_roofline_record_follows_the_reported_schedule():
    """bench.py derives the per-launch roofline from what the library reports (rast_last_schedule): the kernel flavour names the record and
    the batch size sets frames per launch; traffic and limiter come from the committed ncu table of that kernel, scaled to the launch."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.make_workload("spin1080p")
    P = 1920 * 1080
    pass_ms = {"clear": 0.01, "vertex": 0.1, "setup": 0.07, "raster": 3.0, "shade": 7.2}
    sched = {"setup": "k_setup<0,1>", "raster": "k_raster_chunks", "shade": "k_resolve_shade_wt", "batch": "240"}
    rec = bench.roofline_record(wl, "spin1080p", pass_ms, 720, P, 600, 10.0, 720, sched)
    assert rec["kernel"] == "k_resolve_shade_wt" and rec["bound"] == "hbm" and rec["unit"] == "GB/s"
    assert rec["frames_per_launch"] == 240 and abs(rec["avg_launch_ms"] - 7.2 / 3) < 1e-9
    alg = (15 * P + 136 * 600) * 240
    assert abs(rec["algorithmic_bytes_per_launch"] - alg) < 1 and abs(rec["achieved"] - alg / (2.4e-3) / 1e9) < 1e-3
    assert abs(rec["frac"] - rec["achieved"] / rec["peak"]) < 1e-12
    tr, tr_file = bench.ncu_traffic("spin1080p")
    assert tr_file and "k_resolve_shade_wt" in tr, "profiles/r*_traffic.json must hold the kernel the library runs on the headline workload"
    k = tr["k_resolve_shade_wt"]
    assert abs(rec["traffic"] - k["dram_bytes_per_launch"] * 240 / k["frames_per_launch"]) < 1
    assert rec["limiter"]["source"] == tr_file and 0.3 < rec["limiter"]["issue_slot_frac_live"] < 1.0
    # without a schedule (older library): 32 frames per launch and the default kernel names
    rec32 = bench.roofline_record(wl, "spin1080p", pass_ms, 720, P, 600, 10.0, 720, None)
    assert abs(rec32["frames_per_launch"] - 720 / 23) < 1e-9 and rec32["kernel"] == "k_resolve_shade"
