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
    rk = cb["reference_kernels"]  # the reference's own generated kernel text on pre-marshalled arrays, all cores
    assert rk.get("kind") == "reference" and rk["value"] > ref["value"] and rk["scatter_kernel_particles_per_s"] > 0, rk


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


def test_archived_round2_lines_carry_the_contract_keys_and_agree_across_n():
    """The default line as the B200 box printed it at N = 1, 2, 4, 8 (profiles/bench_r2): every key the driver reads, a
    roofline for each leg and each sub-result, and the SAME tally checksum at every N."""
    lines = {}
    for n in (1, 2, 4, 8):
        with open(os.path.join(REPO, "profiles", "bench_r2", "sweep_1b_n%d.json" % n)) as f:
            lines[n] = json.loads(f.readline())
    sys.path.insert(0, REPO)
    import bench

    for n, d in lines.items():
        assert BASE_KEYS | {"roofline", "roofline_photon", "clocks", "tally_checksum", "sub", "detail"} <= set(d), n
        assert d["config"] == bench.SWEEP_CONFIG and d["n_gpus"] == n and d["scaling"] == "strong" and d["dtype"] == "f32"
        assert d["vs_baseline"] is None and d["higher_is_better"] is True and d["gpu_launches"] > 0
        rf = d["roofline"]
        assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and 0.9 < rf["frac"] < 1.0
        rp = d["roofline_photon"]
        assert rp["bound"] == "issue" and abs(rp["frac"] - rp["achieved"] / rp["peak"]) < 1e-9 and rp["hbm_equivalent_frac"] > 1.0
        assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and not d["clocks"]["reasons"]
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        assert e["one_timestep_per_round_trip"]["value"] < e["value"]
        g = d["sub"]["gravity_256k"]
        assert g["roofline"]["bound"] == "fp32" and g["roofline"]["frac"] > 0.70
    assert len({d["tally_checksum"] for d in lines.values()}) == 1
    assert lines[8]["value"] > 7.0 * lines[1]["value"]  # the north star's >= 7x at 8 GPUs on the 1 B-particle sweep
    one = lines[1]
    assert {"value", "unit", "cores", "kind", "sample", "reference_e2e"} <= set(one["cpu_baseline"])
    assert one["cpu_baseline"]["reference_e2e"]["kind"] == "reference"
    assert one["roofline"]["traffic"] and abs(one["roofline"]["traffic"] / (72.0 * 2 ** 30) - 1.0) < 0.01
    for name in ("photon_sphere_16m", "kinematics_64m", "wavelength_64m", "gravity_256k"):
        assert {"value", "roofline", "config"} <= set(one["sub"][name]), name
    assert one["sub"]["kinematics_64m"]["roofline"]["frac"] >= 0.70  # north star: >= 70 % of HBM on the kinematic step
