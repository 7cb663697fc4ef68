"""The committed bench lines (profiles/r1_final_bench_*.json are verbatim stdout of bench.py on a B200) carry every key of the
measurement contract; `bench.py --impl reference` prints a conforming line on the CPU (it times the checker, no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config")


def _load(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.load(f)


def test_headline_line_has_the_contract_keys():
    d = _load("r1_final_bench_50Mrows.json")
    for k in BASE + ("roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["metric"] == "feature_rows_per_sec" and d["unit"] == "rows/s" and d["n_gpus"] == 1 and d["dtype"] == "f64"
    assert d["warmup"] >= 3 and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in r, k
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] >= 3 * d["steps"]
    assert d["clocks"]["sm_mhz"] and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert abs(d["value"] - 50_000_000 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6


def test_multi_gpu_lines_are_whole_job_aggregates():
    one = _load("r1_final_bench_50Mrows.json")
    for name, n in (("r1_final_bench_2gpu_instrument_day.json", 2), ("r1_final_bench_8gpu_instrument_day.json", 8)):
        d = _load(name)
        assert d["n_gpus"] == n and d["scaling"] == "weak"
        assert abs(d["value"] - n * 50_000_000 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
        assert d["value"] > 0.9 * n * one["value"]          # near-linear weak scaling


def test_round2_lines_carry_the_wire_ceiling_and_the_time_shard_record():
    """The shipped round-2 lines: the headline with the e2e record per direction and against the bare-copy ceiling of the same buffers,
    the 8-GPU line with the one-stream-split-by-time sub-record (BASELINE.json configs[4] at its stated size)."""
    d = _load("r2b_bench_50Mrows.json")
    for k in BASE + ("roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert abs(d["value"] - 50_000_000 / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] >= r["algorithmic_bytes"]
    assert abs(sum(r["kernels_ms"].values()) - d["ms_per_step"]) < 0.05 * d["ms_per_step"]      # the three kernels ARE the step
    e = d["e2e"]
    assert 0 < e["value"] < d["value"] and e["h2d_gbs"] > 0 and e["d2h_gbs"] > 0
    c = e["ceiling"]
    assert c["h2d_only_ms"] < c["both_directions_ms"] <= 1.05 * e["ms_per_step"]                 # bare copies are not slower than the call
    assert abs(e["frac_of_copy_ceiling"] - c["both_directions_ms"] / e["ms_per_step"]) < 1e-9
    assert e["norm_only"]["d2h_bytes_per_step"] * 2 == e["d2h_bytes_per_step"] and e["norm_only"]["ms_per_step"] <= e["ms_per_step"]
    d8 = _load("r2b_bench_8gpu.json")
    assert d8["n_gpus"] == 8 and abs(d8["value"] - 8 * 50_000_000 / (d8["ms_per_step"] * 1e-3)) / d8["value"] < 1e-6
    assert d8["value"] > 0.95 * 8 * d["value"]                                                   # weak scaling by instrument-day
    t = d8["timeshard"]
    assert t["rows_total"] == 400_000_000 and t["halo_rows"] == 30416 and t["exchange_ms"] < 0.1 * t["ms_per_step"]
    assert abs(t["value"] - t["rows_total"] / (t["ms_per_step"] * 1e-3)) / t["value"] < 1e-6
    assert t["efficiency_vs_instrument_day_rate"] > 0.9


def test_reference_arm_prints_one_conforming_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-rows", "3000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in BASE + ("impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"] and d["cpu_baseline"]["kind"] in ("reference", "port")
