"""bench.py's reference arm runs on the host cores only: check its JSON contract here (CPU)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the arm must still use every host core
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "shapes/sec" and j["unit"] == "shapes/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["n_gpus"] == 1
    assert j["config"]["workload"].startswith("30k-pt synthetic cloud")
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["value"] == j["value"] and cb["sample"]
    assert cb["cores"] == len(os.sched_getaffinity(0)) == j["config"]["host_threads"]
    e = j["e2e"]
    assert e["value"] == j["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_reference_arm_config0_runs_the_nearest_flow_end_to_end():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                        "--config", "0", "--steps", "1", "--warmup", "0"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    j = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][0])
    assert j["impl"] == "reference" and j["config"]["baseline_config"] == 0
    assert "clock.ply" in j["config"]["workload"] and j["value"] > 0
    assert j["cpu_baseline"]["kind"] == "port" and "whole shapes" in j["cpu_baseline"]["sample"]
