"""CPU checks of bench.py's output contract: the reference arm runs here (it is the CPU oracle on the host cores),
and the archived B200 line of the default workload carries every key the driver reads."""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun exports; the arm must still use every core
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference" and d["metric"] == "particle-steps/s" and d["value"] > 0
    sys.path.insert(0, REPO)
    import bench

    assert d["config"] == bench.SWEEP_CONFIG  # the same static description the GPU arm prints: same_config
    assert d["config"]["workload"] == "sweep_1b" and d["steps"] == 2 and d["warmup"] == 3 and d["scaling"] == "strong"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] == len(os.sched_getaffinity(0)) and cb["sample"]
    ref = cb["reference_e2e"]  # the unmodified reference, end to end (oracle/_ref is built by __graft_entry__.build())
    assert ref.get("kind") == "reference" and ref["value"] > 0 and ref["cores"] == 1, ref


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(REPO, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2"],
                       capture_output=True, text=True, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_archived_round1_line_carries_the_contract_keys():
    with open(os.path.join(REPO, "profiles", "bench_r1", "bench_default.json")) as f:
        d = json.loads(f.readline())
    assert BASE_KEYS | {"roofline", "clocks", "cpu_baseline"} <= set(d)
    assert d["config"]["workload"] == "photon_sphere_16m" and "model" not in d["config"] and d["dtype"] == "f32"
    assert d["vs_baseline"] is None and d["scaling"] == "weak" and d["higher_is_better"] is True and d["gpu_launches"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["traffic"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and not d["clocks"]["reasons"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
