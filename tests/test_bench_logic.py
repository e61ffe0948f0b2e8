"""CPU: host-side logic of bench.py that the driver depends on -- parsing of the nvidia-smi clock samples and their
restriction to the timed region, the config table, the JSON contract keys of the reference arm."""
import datetime
import importlib.util
import os

import pytest

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(_ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _line(ts, sm, smax=1965, reasons=("Not Active",) * 4, power=500.0):
    return "%s, %d, %d, %.2f, %s" % (ts.strftime("%Y/%m/%d %H:%M:%S.%f")[:-3], sm, smax, power, ", ".join(reasons))


def test_clock_samples_are_restricted_to_the_timed_region(bench):
    t = datetime.datetime(2026, 10, 17, 12, 0, 0)
    ms = lambda k: t + datetime.timedelta(milliseconds=k)   # noqa: E731
    out = "\n".join([_line(ms(0), 1200), _line(ms(100), 1500), _line(ms(300), 1965), _line(ms(320), 1950),
                     _line(ms(340), 1965, reasons=("Not Active", "Not Active", "Not Active", "Active")),
                     _line(ms(600), 900), "garbage line", "a, b, c, d, e, f, g, h"])
    clk = bench.ClockSampler.parse(out, ms(295), ms(345))
    assert clk["window"] == "timed region" and clk["samples"] == 3 and clk["samples_total"] == 6
    assert clk["sm_mhz"] == 1965.0 and clk["sm_max_mhz"] == 1965.0
    assert clk["reasons"] == ["sw_power_cap"]
    # no sample inside the window: fall back to everything that was sampled and say so
    clk = bench.ClockSampler.parse(out, ms(1000), ms(1100))
    assert clk["samples"] == 6 and clk["window"].startswith("whole")
    # nothing at all
    clk = bench.ClockSampler.parse("", ms(0), ms(1))
    assert clk["sm_mhz"] is None and clk["samples"] == 0 and clk["reasons"] == []
    # thermal / hardware slowdown flags are reported by name
    hot = _line(ms(10), 1000, reasons=("Active", "Active", "Not Active", "Not Active"))
    assert bench.ClockSampler.parse(hot)["reasons"] == ["hw_slowdown", "hw_thermal_slowdown"]


def test_configs_are_the_baseline_ones(bench):
    cfgs = bench.CONFIGS
    assert {"cfg2", "cfg3", "cfg4", "cfg5"} <= set(cfgs)
    c2 = cfgs["cfg2"]
    assert c2["model"] == "vae" and c2["batch"] == 500 and c2["n_items"] == 50000 and list(c2["dec_dims"]) == [200, 600, 50000]
    assert cfgs["cfg3"]["model"] == "dae" and list(cfgs["cfg3"]["dec_dims"]) == [200, 50000]
    assert cfgs["cfg5"]["n_items"] == 200000
