"""CPU checks of bench.py's contract pieces that need no GPU: the reference arm prints exactly one JSON
line with the agreed keys, the algorithmic-bytes formula of SURVEY.md s8d, the clock summary."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GFLOP/s" and d["higher_is_better"] is True
    assert d["metric"] == "FP64 SpMV GFLOPS" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "banded 10M x 10M" in d["config"]["workload"] and "check vs scalar CSR: pass" in d["config"]["sample"]


def test_algorithmic_bytes_match_survey_table():
    b = _bench()
    # C2 2.12e9 B (13.25 B/nnz), C4 7.43e9 B (8.45 B/nnz)
    assert b.algorithmic_bytes(10_000_000, 10_000_000, 160_000_000, 8) == 2_120_000_004
    c4 = b.algorithmic_bytes(32_768_000, 32_768_000, 879_217_912, 4)
    assert abs(c4 / 879_217_912 - 8.447) < 1e-3
    # the reference's own getB (detail/utils.h:10-14) counts x once per non-zero
    assert b.reference_getB(10, 100, 8) == (10 + 1 + 100) * 4 + (2 * 100 + 10) * 8


def test_clock_summary_windows():
    b = _bench()
    s = b.ClockSampler.__new__(b.ClockSampler)
    s.samples = [(0.5, 300.0, 0x1), (1.5, 1965.0, 0x4), (1.7, 1950.0, 0x0), (3.0, 200.0, 0x1)]
    s.max_mhz = 1965.0
    out = s.summary([(1.0, 2.0)])
    assert out["sm_mhz"] == 1957.5 and out["reasons"] == ["sw_power_cap"] and out["samples"] == 2
    assert s.summary([(10.0, 11.0)])["sampled"] == "whole run"
