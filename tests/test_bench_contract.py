"""bench.py's output contract: exactly one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def run_bench(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--mesh", "medium")
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "mg_cycle_flux_edges_per_s" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    assert d["dtype"] == "f64" and d["data"] == "synthetic" and "workload" in d["config"] and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["cpu_model"] and d["cpu_baseline"]["hardware_threads"] >= 1 and d["scaling"] == "strong"
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["value"] > 0


def test_reference_arm_on_a_per_rank_deck(monkeypatch):
    """a deck that only exists per rank (BASELINE configs[4]): under torchrun rank 0 times its own slab as the bounded
    sample, the other ranks print nothing"""
    monkeypatch.setenv("WORLD_SIZE", "2")
    monkeypatch.setenv("RANK", "0")
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--mesh", "slab_test", "--gpus", "2")
    assert d["impl"] == "reference" and d["value"] > 0 and "x-slab of rank 0 of 2" in d["cpu_baseline"]["sample"]
    assert "generated per rank" in d["config"]["workload"] and d["scaling"] == "strong"
    monkeypatch.setenv("RANK", "1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--mesh", "slab_test", "--gpus", "2"],
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.gpu
def test_b200_arm_line():
    d = run_bench("--steps", "4", "--warmup", "3", "--mesh", "medium", "--cpu-cycles", "1")
    assert BASE_KEYS | {"roofline", "clocks", "mg_cycles_per_s"} <= set(d) and "impl" not in d
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "hbm" and r["unit"] == "GB/s"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and 0 < d["e2e"]["value"] < d["value"]
    assert d["gpu_launches"] > 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["value"] > 0 and {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert d["scaling"] == "strong" and {"frac_fused_L0", "frac_flux_only_L0"} <= set(r)
    assert {"visit_begin", "restrict", "down"} <= set(r["node_kernels"]) and all(v["frac"] > 0 for v in r["node_kernels"].values())
    # the line checks itself against the CPU oracle (one cycle from the initial state) and carries the M6 cycle
    p = d["parity"]
    assert p["checked"] and p["ok"] and p["max_rel_err"] <= 1e-10 and p["validate_count"] == 0
    assert d["m6"]["ms_per_step"] > 0 and "m6" in d["m6"]["workload"]


def test_defaults_are_baseline_configs3():
    """the driver's plain `bench.py --gpus N` must run BASELINE configs[3] (rotor37_8m, strong scaling) at every N"""
    import importlib
    sys.path.insert(0, ROOT)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert 'ap.add_argument("--mesh", default="rotor37_8m"' in src
    assert 'ap.add_argument("--scaling", default="strong"' in src


def test_algorithmic_byte_accounting_matches_the_survey():
    """flux-edge loop 32 E + 120 N per invocation; M6-shaped deck: flux-edge share ~0.68 GB per V-cycle;
    fused stage = flux-edge + time_step (168 N), + residual (120 N) after the last stage; restrict 48 N_f + 216 N_c;
    prolong 148 N_f + 64 N_c; visit prologue copy 80 N + dt 56 N + min 8 N"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    # bench.py redirects fd 1 at import (library banners must not reach stdout): keep this process's stdout
    saved = os.dup(1)
    try:
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
    m6 = [(300000, 930000, 0), (165000, 643000, 0), (111000, 488000, 0), (81000, 377117, 0)]
    assert bench.visits_per_cycle(4) == [0, 1, 2, 3, 2, 1] and bench.visits_per_cycle(1) == [0]
    assert bench.flux_edges_per_cycle(m6) == 3 * (930000 + 2 * 643000 + 2 * 488000 + 377117)
    fb = bench.flux_bytes_per_cycle(m6)
    assert fb == 3 * sum(32 * m6[l][1] + 120 * m6[l][0] for l in [0, 1, 2, 3, 2, 1]) and 0.67e9 < fb < 0.69e9
    sb = bench.rk_stage_bytes_per_cycle(m6)
    assert sb == fb + sum((3 * 168 + 120) * m6[l][0] for l in [0, 1, 2, 3, 2, 1])
    nk = bench.node_kernel_bytes_per_cycle(m6)
    assert nk["visit_begin"] == 144 * (300000 + 2 * 165000 + 2 * 111000 + 81000)
    assert nk["restrict"] == 48 * (300000 + 165000 + 111000) + 216 * (165000 + 111000 + 81000)
    assert nk["down"] == 148 * (300000 + 165000 + 111000) + 64 * (165000 + 111000 + 81000)
    one = [(1000, 3000, 0)]
    assert bench.node_kernel_bytes_per_cycle(one) == {"visit_begin": 144000, "restrict": 0, "down": 0}
