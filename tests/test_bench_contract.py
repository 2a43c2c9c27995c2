"""bench.py contract (CPU side): the reference arm runs here without a GPU and prints ONE JSON line with the keys the
driver reads; the GPU arm's config for the same flags is the same dict (same_config)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mamba_scan_fwd_bwd_GBps_stage1" and d["unit"] == "GB/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "L = 20480" in cb["sample"]
    # the GPU arm builds its config with the same function: identical dicts
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(1, 1)[0]
    assert d["config"]["seqlen"] == 20480 and d["config"]["d_inner"] == 128 and d["config"]["d_state"] == 16


def test_algorithmic_bytes_match_baseline_md():
    """stage 1, B = 1, bf16 -- scan fwd 22 282 240 B, bwd 39 321 600 B."""
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algo_bytes(1) == (22282240, 39321600)
    fwd3, bwd3 = bench.compulsory_bytes(1, 3)
    assert fwd3 < 3 * 22282240 and bwd3 < 3 * 39321600       # z / dout shared by the three directions


import pytest  # noqa: E402


@pytest.mark.gpu
def test_own_arm_prints_one_contract_line(cuda_device):
    """`python bench.py` on one GPU (development flags that skip the slow side legs): ONE JSON line with every key the
    driver and the judge read, internally consistent."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--quick", "--no-vivim", "--steps", "50", "--warmup", "5"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["metric"] == "mamba_scan_fwd_bwd_GBps_stage1" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 50 and d["warmup"] == 5 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["dtype"] == "bf16" and d["vs_baseline"] is None and "impl" not in d
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(1, 1)[0]
    fwd_b, bwd_b = bench.algo_bytes(1)
    # value = algorithmic bytes / measured step time
    assert abs(d["value"] - (fwd_b + bwd_b) / (d["ms_per_step"] * 1e-3) / 1e9) < 1e-6 * d["value"]
    assert 200 < d["value"] < 2000                                   # a B200 runs this step in ~0.12 ms
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["kernel"] == "seg_bwd_kernel"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and abs(r["achieved"] - bwd_b / (r["us_per_launch"] * 1e-6) / 1e9) < 1e-6 * r["achieved"]
    assert r["algorithmic_bytes_per_launch"] == bwd_b and (r["traffic"] is None or r["traffic"] > 0)
    e = d["e2e"]
    assert e["unit"] == "GB/s" and e["h2d_bytes_per_step"] == fwd_b and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["unit"] == "GB/s" and c["cores"] >= 1 and 0 < c["value"] < 10
    assert d["gpu_launches"] == 7 * 50                               # 3 forward + 4 backward kernels per timed step
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert set(d["kernel_us"]) == {"fwd_agg", "fwd_carry", "fwd_main", "bwd_agg", "bwd_carry", "bwd_main", "bwd_cast"}
