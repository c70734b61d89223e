"""The native driver (host/euler3d_b200.cpp = euler3d.cpp's main() over the C-ABI) end to end: reads a deck, runs
N cycles, validates with the reference's -v criterion against the solution files, writes --output-variables."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "mg-cfd-app-op2_b200", "euler3d_b200")


def make_deck(tmp_path, meshgen, golden):
    mesh = meshgen.make_multigrid("small")
    meshgen.write_deck(str(tmp_path), mesh)
    g = golden("small_cycles10.npz")
    for l in range(len(mesh["levels"])):
        meshgen.write_solution(str(tmp_path), l, 10, g[f"var_L{l}"])
    return mesh, g


@pytest.mark.parametrize("extra", [[], ["--loopwise"], ["--loopwise", "-b", "--variant", "colour"], ["--exact"],
                                   ["-b"], ["-b", "--gpus", "2", "--same-device", "--exact"],
                                   ["--gpus", "2", "--same-device"], ["--gpus", "4", "--same-device", "--exact"]])
def test_driver_validates(tmp_path, meshgen, golden, extra):
    mesh, g = make_deck(tmp_path, meshgen, golden)
    out = os.path.join(str(tmp_path), "out.")
    cmd = [EXE, "-i", "input.dat", "-d", str(tmp_path), "-g", "10", "-v", "--output-variables", "-o", out] + extra
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Validation passed" in p.stdout and "Max total runtime" in p.stdout
    for l in range(len(mesh["levels"])):
        got = meshgen.read_container(f"{out}variables.L{l}.cycles=10.mgb")[f"p_variables_result_L{l}"]
        ref = g[f"var_L{l}"]
        if "--exact" in extra:
            assert np.array_equal(got, ref)
        else:
            assert (np.abs(got - ref).max(axis=0) <= 1e-10 * np.abs(ref).max(axis=0)).all()


def test_driver_reports_failed_validation(tmp_path, meshgen, golden):
    mesh, g = make_deck(tmp_path, meshgen, golden)
    meshgen.write_solution(str(tmp_path), 0, 10, g["var_L0"] * (1 + 1e-5))
    p = subprocess.run([EXE, "-i", "input.dat", "-d", str(tmp_path), "-g", "10", "-v"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "Validation failed" in p.stdout and "Validation passed" not in p.stdout


def test_driver_argument_errors(tmp_path):
    p = subprocess.run([EXE], capture_output=True, text=True)
    assert p.returncode == 1 and "input_file not set" in p.stdout
    p = subprocess.run([EXE, "-i", "missing.dat", "-d", str(tmp_path)], capture_output=True, text=True)
    assert p.returncode == 1 and "Could not open input file" in p.stderr


def test_driver_hdf5_deck_end_to_end(tmp_path, meshgen, golden):
    """the reference's own file formats end to end: HDF5 level files in, -v against solution.variables.L<l>.cycles=<g>.h5
    (euler3d.cpp:314-335), --output-variables written as HDF5 (op_fetch_data_hdf5_file, euler3d.cpp:740-770)"""
    mesh = meshgen.make_multigrid("small")
    meshgen.write_deck(str(tmp_path), mesh, fmt="h5")
    g = golden("small_cycles10.npz")
    for l in range(len(mesh["levels"])):
        meshgen.write_solution(str(tmp_path), l, 10, g[f"var_L{l}"], fmt="h5")
    out = os.path.join(str(tmp_path), "out.")
    cmd = [EXE, "-i", "input.dat", "-d", str(tmp_path), "-g", "10", "-v", "--output-variables", "-o", out, "--exact"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Validation passed" in p.stdout
    for l in range(len(mesh["levels"])):
        got = meshgen.read_h5(f"{out}variables.L{l}.cycles=10.h5")[f"p_variables_result_L{l}"]
        assert np.array_equal(got, g[f"var_L{l}"])


@pytest.mark.parametrize("extra", [[], ["--gpus", "2", "--same-device"]])
def test_driver_periodic_flow_dumps_on_the_default_path(tmp_path, meshgen, golden, extra):
    """-I <n> (euler3d.cpp:552-571) on the device-driven schedule: the level-0 flow is written every n cycles and the run
    still ends on the 10-cycle reference solution (round 1 honoured -I only with --loopwise)"""
    mesh, g = make_deck(tmp_path, meshgen, golden)
    out = os.path.join(str(tmp_path), "out.")
    cmd = [EXE, "-i", "input.dat", "-d", str(tmp_path), "-g", "10", "-v", "-I", "4", "--exact", "-o", out] + extra
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "Validation passed" in p.stdout
    dumps = sorted(f for f in os.listdir(str(tmp_path)) if f.startswith("out.variables.L0.cycle="))
    assert dumps == ["out.variables.L0.cycle=4.mgb", "out.variables.L0.cycle=8.mgb"]
    v4 = meshgen.read_container(os.path.join(str(tmp_path), dumps[0]))["p_variables"]
    v8 = meshgen.read_container(os.path.join(str(tmp_path), dumps[1]))["p_variables"]
    assert v4.shape == g["var_L0"].shape and np.isfinite(v4).all() and not np.array_equal(v4, v8)


@pytest.mark.parametrize("extra", [[], ["--gpus", "2", "--same-device"]])
def test_driver_perf_csvs_with_timers(tmp_path, meshgen, golden, extra):
    """--timers: device time per call site (inside graph replay on one GPU) lands in the reference's CSVs -- the flux rows of
    P=<r>.PerfData.csv (io.h:296-417) and op2_performance_data.csv (op_timings_to_csv, euler3d.cpp:651-656), which
    run-scripts/aggregate-output-data.py:41,71 reads for `nranks` and the per-loop times"""
    make_deck(tmp_path, meshgen, golden)
    out = os.path.join(str(tmp_path), "run.")
    p = subprocess.run([EXE, "-i", "input.dat", "-d", str(tmp_path), "-g", "10", "-v", "--timers", "-o", out] + extra,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "Validation passed" in p.stdout, p.stdout + p.stderr
    n_ranks = 2 if extra else 1
    rows = [l.split(",") for l in open(out + "op2_performance_data.csv").read().splitlines()]
    assert rows[0] == ["rank", "thread", "nranks", "nthreads", "count", "total time", "plan time", "mpi time", "GB used", "GB total", "kernel name"]
    stage = [r for r in rows[1:] if r[-1] == "rk_stage"]
    assert len(stage) == n_ranks and all(int(r[2]) == n_ranks and int(r[4]) > 0 and float(r[5]) > 0 for r in stage)
    for r in range(n_ranks):
        perf = [l.split(",") for l in open(f"{out}P={r}.PerfData.csv").read().splitlines()]
        assert perf[0] == ["rank", "partitioner", "kernel", "level", "computeTime", "syncTime", "iters"]
        assert all(row[2] == "compute_flux_edge_kernel" and float(row[4]) > 0 and int(row[6]) > 0 for row in perf[1:])
