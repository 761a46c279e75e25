"""The committed bench lines (written by bench.py on a B200, profiles/) carry every key of the bench contract."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


import pytest


@pytest.mark.parametrize("name", ["r1_bench_line.json", "r2_bench_c2.json", "r2_bench_c1.json", "r2_bench_c4.json", "r2_bench_c5.json"])
def test_our_arm_line_has_the_contract_keys(name):
    d = _line(name)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["metric"] == "separator frames/sec" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]                 # copies inside the timed region cost something
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    r = d["roofline"]
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(r)
    assert r["bound"] in ("hbm", "tensor") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    if c is not None:               # the extra-workload lines (c4, c5) were taken with --no-cpu-baseline
        assert set(("value", "unit", "cores", "kind", "sample")) <= set(c) and c["kind"] in ("port", "reference")
    else:
        assert name in ("r2_bench_c4.json", "r2_bench_c5.json")
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
    if name.startswith("r2"):      # round 2: parity evidence, HBM view of the roofline and the reference on the same GPU in the line
        assert d["parity"]["ok"] is True and d["parity"]["rel_l2_vs_fp32_path"] < d["parity"]["tolerance"] <= 1e-3
        assert "hbm_frac" in r and "whole_step" in r and "kernel_ms" in d and "reference_gpu" in d
        assert not any(x in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown") for x in d["clocks"]["reasons"])


def test_two_gpu_line_reports_the_whole_job():
    d, d1 = _line("r2_bench_n2.json"), _line("r2_bench_c2.json")
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and d["config"]["global_batch"] == 2 * d1["config"]["global_batch"]
    assert 1.8 * d1["value"] < d["value"] < 2.05 * d1["value"]


@pytest.mark.parametrize("name,kind", [("r1_bench_reference_line.json", "port"), ("r2_bench_reference_cpu.json", "reference")])
def test_reference_arm_line_has_the_contract_keys(name, kind):
    d = _line(name)
    assert d["impl"] == "reference" and d["metric"] == "separator frames/sec" and d["unit"] == "frames/s"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] and d["cpu_baseline"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1
