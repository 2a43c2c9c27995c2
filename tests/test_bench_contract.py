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
