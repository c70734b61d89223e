"""Pins the CPU oracle.  The reference tree holds no golden vectors or unit tests
(SURVEY.md 4.1), so the oracle is pinned three ways:
  1. tests/golden/*.npz were produced by oracle/_ref -- the reference's own elemental-kernel headers
     compiled in place -- and the plain-C port must reproduce them bit for bit (runs everywhere);
  2. where oracle/_ref exists (build container; it also travels to the GPU box prebuilt), the port and
     _ref are compared bit for bit on every loop and on whole multigrid runs;
  3. the reference's own -v criterion (validation.h:46-100) is evaluated on perturbed solutions.
"""
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT, mesh0


def _gen_golden():
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(ROOT, "oracle", "gen_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_constants(oracle_port, golden):
    c = golden("consts.npz")["consts"]
    assert np.array_equal(oracle_port.consts, c)
    assert c[0] == float(np.float32(0.2))             # Q1: double(0.2f), not 0.2
    assert c[0] != 0.2
    assert np.allclose(c[1:6], [1.4, 1.4 * 1.2, 0.0, 0.0, 1.4 * 0.5 * 1.44 + 2.5])   # Q11


@pytest.mark.parametrize("name,cycles", [("tiny", 3), ("small", 10)])
def test_port_reproduces_golden_cycles(oracle_port, meshgen, golden, name, cycles):
    g = golden(f"{name}_cycles{cycles}.npz")
    got = _gen_golden().cycle_vectors(oracle_port, mesh0(meshgen, name), cycles)
    assert set(got) == set(g)
    for k in g:
        assert np.array_equal(got[k], g[k]), k


def test_port_reproduces_golden_loops(oracle_port, meshgen, golden):
    g = golden("tiny_loops.npz")
    got = _gen_golden().loop_vectors(oracle_port, mesh0(meshgen, "tiny"))
    assert set(got) == set(g)
    for k in g:
        assert np.array_equal(got[k], g[k]), k


@pytest.mark.parametrize("name,cycles", [("tiny", 3), ("small", 10), ("medium", 2)])
def test_port_equals_reference_headers(oracle_port, oracle_ref, meshgen, name, cycles):
    gg = _gen_golden()
    lev = mesh0(meshgen, name)
    a, b = gg.cycle_vectors(oracle_ref, lev, cycles), gg.cycle_vectors(oracle_port, lev, cycles)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    a, b = gg.loop_vectors(oracle_ref, lev, seed=77), gg.loop_vectors(oracle_port, lev, seed=77)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_golden_is_current(oracle_ref, meshgen, golden):
    """the committed fixtures are what oracle/_ref produces today (guards against stale goldens)"""
    g = golden("tiny_cycles3.npz")
    got = _gen_golden().cycle_vectors(oracle_ref, mesh0(meshgen, "tiny"), 3)
    for k in g:
        assert np.array_equal(got[k], g[k]), k


def test_boundary_dispatch_q2(oracle_port):
    """flux.h:29-37: g<=2 (incl. negatives) pressure wall, 3..7 far field, >7 no-op"""
    var = np.array([[1.3, 1.5, 0.02, -0.01, 3.4]])
    w = np.array([[1e-3, -2e-3, 5e-4]])
    out = {}
    for grp in (-1, 0, 2, 3, 4, 7, 8, 9):
        f = np.zeros((1, 5))
        oracle_port.compute_bnd_node_flux(np.array([grp], np.int32), w, np.array([0], np.int32), var, f)
        out[grp] = f.copy()
    assert np.array_equal(out[-1], out[0]) and np.array_equal(out[0], out[2])
    assert out[0][0, 0] == 0 and out[0][0, 4] == 0 and out[0][0, 1] != 0
    assert np.array_equal(out[3], out[4]) and np.array_equal(out[4], out[7]) and out[3][0, 0] != 0
    assert not out[8].any() and not out[9].any()


def test_validation_criterion_q13(oracle_port):
    ref = np.array([[1.4, 1.68, 0.0, -1e-9, 3.5]] * 4)
    t = ref.copy()
    assert oracle_port.validate_count(t, ref) == 0
    t[0, 0] *= 1 + 0.9e-7       # inside 1e-7 relative
    t[1, 1] *= 1 + 1.1e-7       # outside
    t[2, 2] = 2e-19             # inside the 3e-19 absolute floor
    t[3, 2] = 4e-19             # outside
    assert oracle_port.validate_count(t, ref) == 2


def test_restrict_keeps_childless_q8(oracle_port):
    mg = np.array([0, 0, 2], np.int32)            # coarse node 1 has no child
    var = np.arange(15, dtype=np.float64).reshape(3, 5)
    above = np.full((3, 5), 7.0)
    sc = np.zeros((3, 2), np.int32)
    oracle_port.up_pre(mg, above, sc); oracle_port.up(mg, var, above, sc); oracle_port.up_post(above, sc)
    assert np.array_equal(above[0], (var[0] + var[1]) * 0.5)
    assert np.array_equal(above[1], np.full(5, 7.0))
    assert np.array_equal(above[2], var[2])
    assert list(sc[:, 0]) == [2, 0, 1]


def test_openmp_mode_matches_seq_within_rounding(orc_mod, meshgen):
    """the multi-threaded baseline mode computes the same thing (increment order differs)"""
    o = orc_mod.Oracle("port_fast")
    lev = mesh0(meshgen, "small")
    a = o.make_state(lev); a.init(); a.run(3)
    assert o.set_threads(4) >= 1
    b = o.make_state(lev); b.init(); b.run(3)
    o.set_threads(1)
    for x, y in zip(a.levels, b.levels):
        assert np.abs(x["var"] - y["var"]).max() <= 1e-12
