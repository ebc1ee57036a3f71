"""The evidence tools keep working on the committed ncu launch lists (they are what turns a
`gpurun_out/*.csv` into the tables under profiles/)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FWD = os.path.join(ROOT, "profiles", "r02_unet_forward_launches.csv")


def _run(*args):
    return subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, check=True).stdout


def test_summarize_ncu_on_the_forward_launch_list():
    out = _run("tools/summarize_ncu.py", FWD)
    head = out.splitlines()[0]
    assert "356 launches" in head
    rows = {l.split()[0]: l.split() for l in out.splitlines()[2:] if l.strip()}
    conv_ms = sum(float(v[-3]) for k, v in rows.items() if k.startswith(("conv_", "splitk_")))
    total_ms = float(head.split("launches,")[1].split("ms")[0])
    # the tcgen05 convs are ~80 % of a forward, as the bench line reports in situ
    assert 0.75 < conv_ms / total_ms < 0.88


def test_layer_table_enumerates_every_conv_layer():
    out = _run("tools/layer_table.py", FWD)
    assert "117 conv launches" in out.splitlines()[0]
    assert "conv_halo_kernel<1, 256, 3, 9>" in out and "stem" in out


def test_conv_traffic_matches_the_committed_json(tmp_path):
    dst = tmp_path / "t.json"
    _run("tools/conv_traffic.py", FWD, str(dst), "test")
    got = json.load(open(dst))
    ref = json.load(open(os.path.join(ROOT, "profiles", "r02_conv_traffic.json")))
    assert got["launches"] == ref["launches"] == 117
    assert got["avg_traffic_bytes_per_launch"] == ref["avg_traffic_bytes_per_launch"]
    # DRAM traffic of the convs stays at the algorithmic level: < 130 MB per launch on average
    assert 1.0e8 < got["avg_traffic_bytes_per_launch"] < 1.3e8
