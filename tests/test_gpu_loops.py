"""GPU parity, loop by loop: every op_par_loop call site of the cycle through the
C-ABI against the CPU oracle on the same inputs, and against the committed golden vectors that
oracle/_ref (the reference's own headers) produced.

Tolerances (BASELINE.json north_star: per-variable state within 1e-10 relative):
  * node loops and the exact-arithmetic owner flux are compared BIT FOR BIT;
  * fast-arithmetic flux variants: normwise relative <= 1e-12 on the flux array.
"""
import numpy as np
import pytest

from conftest import mesh0

pytestmark = pytest.mark.gpu

VARIANTS = ["owner", "emit", "gather", "colour", "atomic"]
FLUX_TOL = 1e-12


def physical_state(n, ff, seed):
    rng = np.random.default_rng(seed)
    u = ff[None, :] * (1.0 + rng.uniform(-0.05, 0.05, size=(n, 5)))
    u[:, 2:4] = rng.uniform(-0.05, 0.05, size=(n, 2))
    return np.ascontiguousarray(u)


def normwise(a, b):
    return np.abs(a - b).max(axis=0) / np.maximum(np.abs(b).max(axis=0), 1e-300)


@pytest.fixture(scope="module")
def tiny(meshgen):
    return meshgen.make_multigrid("tiny")


@pytest.fixture(scope="module")
def small(meshgen):
    return meshgen.make_multigrid("small")


def make_gpu(pkg, mesh, **kw):
    return pkg.MGCFD(mesh["levels"], base_array_index=mesh["base_array_index"], **kw)


def test_init_state_bit_exact(pkg, tiny, golden):
    g = golden("tiny_loops.npz")
    with make_gpu(pkg, tiny) as gpu:
        assert np.array_equal(gpu.fetch(0, "volumes"), g["init_vol"])
        assert np.array_equal(gpu.fetch(0, "edge_weights"), g["init_ewt"])
        assert np.array_equal(gpu.fetch(0, "bnd_node_weights"), g["init_bwt"])
        ff = np.array(list(gpu.consts.ff_variable))
        assert np.array_equal(gpu.fetch(0, "variables"), np.tile(ff, (g["init_vol"].shape[0], 1)))
        assert not gpu.fetch(0, "fluxes").any()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("exact", [False, True])
def test_flux_edge_golden(pkg, tiny, golden, variant, exact):
    if variant == "emit" and exact:
        pytest.skip("the emit variant is fast-arithmetic only")
    g = golden("tiny_loops.npz")
    with make_gpu(pkg, tiny, flux_variant=variant, exact_arith=exact) as gpu:
        gpu.set(0, "variables", g["in_var"])
        gpu.set(0, "fluxes", g["in_flux"])
        gpu.compute_flux_edge(0)
        got = gpu.fetch(0, "fluxes")
    if exact and variant in ("owner", "gather"):
        assert np.array_equal(got, g["flux_edge"])
    else:
        # compare the increment, not in_flux + increment, so that the tolerance is not diluted
        inc, ref = got - g["in_flux"], g["flux_edge"] - g["in_flux"]
        assert (normwise(inc, ref) <= 1e-8).all()   # all the difference in_flux+inc-in_flux can resolve
        assert (normwise(got, g["flux_edge"]) <= FLUX_TOL).all()


@pytest.mark.parametrize("variant", VARIANTS)
@pytest.mark.parametrize("exact", [False, True])
def test_flux_edge_from_zero_vs_oracle(pkg, small, meshgen, oracle_port, variant, exact):
    """flux starts at zero (the state every flux loop sees in the cycle): small mesh, live oracle."""
    if variant == "emit" and exact:
        pytest.skip("the emit variant is fast-arithmetic only")
    lev = mesh0(meshgen, "small")
    run = oracle_port.make_state(lev)
    run.init()
    L0 = run.levels[0]
    var = physical_state(L0["var"].shape[0], oracle_port.ff_variable, 99)
    ref = np.zeros_like(var)
    oracle_port.compute_flux_edge(L0["e2n"], var, L0["ewt"], ref)
    with make_gpu(pkg, small, flux_variant=variant, exact_arith=exact) as gpu:
        gpu.set(0, "variables", var)
        gpu.compute_flux_edge(0)
        got = gpu.fetch(0, "fluxes")
        if exact and variant in ("owner", "gather"):
            assert np.array_equal(got, ref)
        else:
            assert (normwise(got, ref) <= FLUX_TOL).all(), normwise(got, ref)
        # conservation: interior-edge increments cancel pairwise
        assert (np.abs(got.sum(axis=0)) <= 1e-9 * np.abs(got).sum(axis=0) + 1e-300).all()
        # a second loop accumulates (OP_INC), it does not overwrite
        gpu.compute_flux_edge(0)
        got2 = gpu.fetch(0, "fluxes")
        assert (normwise(got2, 2 * ref) <= FLUX_TOL).all()


@pytest.mark.parametrize("variant", VARIANTS)
def test_unstructured_stream(pkg, tiny, golden, variant):
    g = golden("tiny_loops.npz")
    with make_gpu(pkg, tiny, flux_variant=variant) as gpu:
        gpu.set(0, "variables", g["in_var"])
        gpu.set(0, "dummy_fluxes", g["in_flux"])
        gpu.unstructured_stream(0)
        got = gpu.fetch(0, "dummy_fluxes")
    assert (normwise(got, g["ustream"]) <= 1e-13).all()


@pytest.mark.parametrize("exact", [False, True])
def test_bnd_flux_golden(pkg, tiny, golden, exact):
    g = golden("tiny_loops.npz")
    with make_gpu(pkg, tiny, exact_arith=exact) as gpu:
        gpu.set(0, "variables", g["in_var"])
        gpu.set(0, "fluxes", g["in_flux"])
        gpu.compute_bnd_node_flux(0)
        got = gpu.fetch(0, "fluxes")
    if exact:
        assert np.array_equal(got, g["bnd_flux"])
    else:
        assert (normwise(got, g["bnd_flux"]) <= 1e-13).all()


def test_node_loops_bit_exact(pkg, tiny, golden):
    g = golden("tiny_loops.npz")
    with make_gpu(pkg, tiny) as gpu:
        gpu.set(0, "variables", g["in_var"])
        gpu.calculate_dt(0)
        assert np.array_equal(gpu.fetch(0, "step_factors"), g["dt"])
        m = gpu.get_min_dt(0)
        assert m == g["min_dt"][0]
        assert gpu.get_min_dt(0, start=m / 2) == m / 2          # OP_MIN keeps a smaller incoming value
        gpu.compute_step_factor(0, m)
        assert np.array_equal(gpu.fetch(0, "step_factors"), g["step_factor"])
        for rk in range(3):
            gpu.set(0, "old_variables", g["in_old"])
            gpu.set(0, "fluxes", g["flux_edge"])
            gpu.time_step(0, rk)
            assert np.array_equal(gpu.fetch(0, "variables"), g[f"time_step_rk{rk}"])
            assert not gpu.fetch(0, "fluxes").any()              # Q7: time_step zeroes the fluxes
        gpu.set(0, "variables", g["in_var"])
        gpu.set(0, "old_variables", g["in_old"])
        gpu.residual(0)
        assert np.array_equal(gpu.fetch(0, "residuals"), g["residual"])
        rms = gpu.calc_rms(0)
        assert abs(rms - g["rms"][0]) <= 1e-13 * g["rms"][0]
        assert gpu.count_bad_vals(0) == 0
        bad = g["in_var"].copy()
        bad[3, 1] = np.nan
        bad[5, 4] = np.inf
        gpu.set(0, "variables", bad)
        assert gpu.count_bad_vals(0) == 2
        gpu.copy_double(0)
        assert np.array_equal(gpu.fetch(0, "old_variables"), bad, equal_nan=True)


def test_restrict_prolong_bit_exact(pkg, tiny, golden):
    g = golden("tiny_loops.npz")
    with make_gpu(pkg, tiny) as gpu:
        gpu.set(0, "variables", g["in_var"])
        gpu.set(1, "variables", g["in_var_above"])
        gpu.up_pre(1)
        gpu.up(1)
        gpu.up_post(1)
        assert np.array_equal(gpu.fetch(1, "up_scratch"), g["restrict_count"])
        assert (g["restrict_count"] == 0).any(), "fixture must exercise childless coarse nodes (Q8)"
        assert np.array_equal(gpu.fetch(1, "variables"), g["restrict"])
        gpu.set(0, "variables", g["in_var"])
        gpu.set(0, "residuals", g["residual"])
        gpu.set(1, "residuals", g["in_res_above"])
        gpu.down(0)
        assert np.array_equal(gpu.fetch(0, "variables"), g["prolong"])


def test_restrict_of_constant_is_constant(pkg, small):
    with make_gpu(pkg, small) as gpu:
        n0, n1 = gpu.sizes[0][0], gpu.sizes[1][0]
        gpu.set(0, "variables", np.full((n0, 5), 3.25))
        gpu.set(1, "variables", np.full((n1, 5), 3.25))
        gpu.up_pre(1); gpu.up(1); gpu.up_post(1)
        assert np.array_equal(gpu.fetch(1, "variables"), np.full((n1, 5), 3.25))
