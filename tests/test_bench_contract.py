"""CPU-side guard of the bench contract: the committed line of the last GPU run (profiles/r01_bench_line.json, written by
`python bench.py` on a B200) carries every key the driver and the judge read, with consistent values; bench.py's command
line accepts what the driver passes.  Nothing here runs the hot path (no GPU in this container)."""
import ast
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINE = os.path.join(ROOT, "profiles", "r01_bench_line.json")


@pytest.fixture(scope="module")
def line():
    if not os.path.exists(LINE):
        pytest.skip("no committed bench line")
    return json.loads(open(LINE).read().strip().splitlines()[-1])


def test_base_contract_keys(line):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in line, k
    baseline = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["metric"].split(" (")[0] == baseline["metric"].split(" (")[0] == "frames/sec"
    assert line["unit"] == "frames/s" and line["higher_is_better"] is True and line["vs_baseline"] is None  # nothing published
    assert line["data"] == "synthetic" and line["dtype"] == "f32" and "workload" in line["config"] and "model" not in line["config"]
    assert line["warmup"] >= 3 and line["n_gpus"] == 1
    assert abs(line["value"] - 1e3 / line["ms_per_step"]) <= 1e-6 * line["value"]
    assert line["gpu_launches"] > 0


def test_e2e_roofline_cpu_baseline_clocks(line):
    e = line["e2e"]
    assert e["unit"] == line["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != line["value"]  # measured on its own, through the host-buffer API
    r = line["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) <= 1e-9 and r["peak"] > 0
    assert r["traffic"] is None or r["traffic"] > 0
    # the dominant kernel is the one with the largest share of the step
    assert r["kernel_ms"] == max(k["kernel_ms"] for k in line["roofline_kernels"])
    assert sum(k["share_of_step"] for k in line["roofline_kernels"]) < 1.0
    c = line["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == line["unit"] and c["sample"]
    k = line["clocks"]
    assert k["sm_mhz"] > 0.9 * k["sm_max_mhz"]
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_bench_command_line_matches_the_driver():
    """--gpus / --steps / --warmup / --impl reference, defaults N = 1 and W >= 3"""
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    args = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument" and node.args:
            name = node.args[0].value
            args[name] = {kw.arg: getattr(kw.value, "value", None) for kw in node.keywords}
    assert args["--gpus"]["default"] == 1 and args["--warmup"]["default"] >= 3 and args["--steps"]["default"] > 0
    assert "--impl" in args
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert '"impl": "reference"' in src and "/root/reference" not in src  # the reference tree does not exist on the GPU box
