"""bench.py's host-side logic without a GPU: the reference arm's JSON line (it drives the unmodified reference binary on a small sample
of the workload) and the roofline accounting (SURVEY 8(d) bytes x units counted by the library / live kernel time)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "3", "--warmup", "3",
                        "--reads-per-step", "40000", "--genome", "2000000", "--bf-shift", "24"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout                       # exactly one JSON line on stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "events/s" and d["scaling"] == "strong"
    assert d["steps"] == 3 and d["warmup"] == 3 and d["n_gpus"] == 1 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["config"]["workload"] == "cfg2" and d["config"]["reads_per_step"] == 40000 and d["config"]["k"] == 31
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "events/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # every rank but 0 leaves without work or output (the driver launches the arm under torchrun at N > 1)
    r2 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True, timeout=60,
                        cwd=ROOT, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"))
    assert r2.returncode == 0 and r2.stdout.strip() == ""


def test_both_arms_describe_the_same_config():
    sys.path.insert(0, ROOT)
    code = ("import sys, json, argparse; sys.argv=['bench.py']; import bench; "
            "a=argparse.Namespace(reads_per_step=16000000, bf_shift=37, genome=3000000000); "
            "sys.stderr.write(json.dumps([bench.base_config(a, 1), bench.base_config(a, 8)]))")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0, r.stderr[-1000:]
    c1, c8 = json.loads(r.stderr.strip().splitlines()[-1])
    assert c1 == c8 and c1["workload"] == "cfg2" and "model" not in c1     # the config names the workload, identical at every N and in both arms


def test_roofline_accounting():
    code = r"""
import sys, json
sys.argv = ['bench.py']
import bench
prof = {"zone_probe": [1000.0, 20, 30_000_000_000], "part_scatter": [400.0, 20, 30_000_000_000], "group_insert": [300.0, 25, 5_000_000_000],
        "journal(sort+seg)": [2000.0, 20, 1], "host:cudaMalloc": [5000.0, 3, 0]}
dom, per, pi = bench.roofline_objects(prof, 2000.0, 6545.0, "test")
sys.stderr.write(json.dumps({"dom": dom, "per": {k: v["frac"] for k, v in per.items()}, "pi": pi}))
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT, timeout=120)
    assert r.returncode == 0, r.stderr[-1000:]
    d = json.loads(r.stderr.strip().splitlines()[-1])
    # the dominant KERNEL with a byte model - not the longest scope (journal) and not host time
    assert d["dom"]["kernel"] == "zone_probe" and d["dom"]["bound"] == "hbm" and d["dom"]["unit"] == "GB/s"
    assert abs(d["dom"]["achieved"] - 24.0 * 30e9 / 1.0 / 1e9) < 1e-6 and abs(d["dom"]["frac"] - 720.0 / 6545.0) < 1e-9
    assert abs(d["dom"]["algorithmic_bytes_per_launch"] - 24.0 * 30e9 / 20) < 1 and abs(d["dom"]["share_of_step"] - 0.5) < 1e-9
    assert set(d["per"]) == {"zone_probe", "part_scatter", "group_insert"}
    assert abs(d["pi"]["achieved"] - (8.31 + 24.0) * 30e9 / 1.4 / 1e9) < 1e-6
