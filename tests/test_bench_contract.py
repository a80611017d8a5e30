"""The bench line is a contract with the driver: check the committed line of the last GPU run (profiles/) and the
argument surface of bench.py without needing a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    txt = open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1]
    return json.loads(txt)


def test_committed_bench_line_has_the_contract_keys():
    d = _line("r1_bench_bf128_l0_n1.json")
    for k, typ in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int),
                   ("ms_per_step", float), ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str),
                   ("config", dict), ("clocks", dict), ("e2e", dict), ("gpu_launches", int), ("roofline", dict),
                   ("cpu_baseline", dict)):
        assert isinstance(d[k], typ), k
    assert "vs_baseline" in d and d["vs_baseline"] is None  # BASELINE.md publishes no number for this metric
    # the archived line was printed before bench.py switched to BASELINE.json's full spelling of the metric
    assert json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"].startswith(d["metric"])
    assert d["warmup"] >= 3 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "bf128_l0" and "l2" in d["config"]
    assert d["gpu_launches"] > 0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] <= d["value"] * 1.01
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] > 0
    assert 0 < r["issue"]["frac"] < 1
    c = d["cpu_baseline"]
    assert c["kind"] in ("port", "reference") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    cl = d["clocks"]
    assert cl["sm_mhz"] > 0 and cl["sm_max_mhz"] >= cl["sm_mhz"] and isinstance(cl["reasons"], list)
    assert not set(cl["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_committed_reference_arm_line():
    d = _line("r1_bench_reference_arm_n1.json")
    assert d["impl"] == "reference" and d["metric"] == _line("r1_bench_bf128_l0_n1.json")["metric"]
    assert d["config"]["workload"] == "bf128_l0"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"]


def test_bench_cli_surface():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"], capture_output=True, text=True)
    assert out.returncode == 0
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert flag in out.stdout


def test_bench_prints_baseline_metric_verbatim():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.METRIC == json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]


def test_round2_lines_same_config_and_parity():
    """Round 2: the two arms print the SAME config object (the driver compares them), the CPU arm is the reference's own
    code with a reproducible thread count, and the GPU line carries the in-run parity object."""
    g = _line("r2_bench_bf128_l0_n1.json")
    r = _line("r2_bench_reference_arm_n1.json")
    assert g["config"] == r["config"] and g["metric"] == r["metric"] and g["unit"] == r["unit"]
    assert r["impl"] == "reference" and r["cpu_baseline"]["kind"] == "reference" and r["cpu_baseline"]["steps"] >= 3
    c = g["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] == r["cpu_baseline"]["cores"]
    assert abs(c["value"] - r["value"]) / r["value"] < 0.10  # reproducible across runs / boxes (round 1: 4.3x apart)
    p = g["parity"]
    assert p["against"] == "reference" and p["index_mismatches"] == 0 and p["cost_mismatch_frac"] <= 1e-5
    assert p["pixel_candidates"] > 1e6
    assert g["cfg1_full"]["parity"]["index_mismatches"] == 0
    assert g["e2e"]["value"] > 0 and g["roofline"]["frac"] > 0.08 and g["gpu_launches"] > 0
    c2f = _line("r2_bench_c2f5_n1.json")
    assert c2f["config"]["workload"] == "c2f5" and c2f["roofline"]["kernel"] == "pingPongKernel"
    assert c2f["e2e"]["ms_per_step"] >= c2f["ms_per_step"] and c2f["parity"]["fraction"] >= 0.999
