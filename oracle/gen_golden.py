"""Regenerate tests/golden/*.npz from oracle/_ref (the reference's own headers compiled in place).

Run in the build container (where /root/reference exists):  python oracle/gen_golden.py
The fixtures pin (a) the plain-C port on machines without /root/reference and (b) the CUDA
path's parity tests.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import orc  # noqa: E402


def _meshgen():
    spec = importlib.util.spec_from_file_location("mgcfd_meshgen", os.path.join(ROOT, "mg-cfd-app-op2_b200", "meshgen.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def physical_state(n, ff, seed):
    """SURVEY.md 8d value distribution: ff_variable*(1+U(-0.05,0.05)), transverse momenta U(-0.05,0.05)."""
    rng = np.random.default_rng(seed)
    u = ff[None, :] * (1.0 + rng.uniform(-0.05, 0.05, size=(n, 5)))
    u[:, 2:4] = rng.uniform(-0.05, 0.05, size=(n, 2))
    return np.ascontiguousarray(u)


def loop_vectors(o, mesh0, seed=1234):
    """Inputs and outputs of every op_par_loop call site on level 0/1 of a mesh, from a perturbed state."""
    run = o.make_state(mesh0)
    run.init()
    L0, L1 = run.levels[0], run.levels[1]
    ff = o.ff_variable
    out = {}
    var = physical_state(L0["var"].shape[0], ff, seed)
    out["in_var"] = var.copy()
    out["init_vol"] = L0["vol"].copy()
    out["init_ewt"] = L0["ewt"].copy()
    out["init_bwt"] = L0["bwt"].copy()
    rng = np.random.default_rng(seed + 1)
    flux = rng.uniform(-1e-3, 1e-3, size=var.shape)
    out["in_flux"] = flux.copy()
    f = flux.copy(); o.compute_flux_edge(L0["e2n"], var, L0["ewt"], f); out["flux_edge"] = f
    f = flux.copy(); o.compute_bnd_node_flux(L0["bgroup"], L0["bwt"], L0["b2n"], var, f); out["bnd_flux"] = f
    f = flux.copy(); o.unstructured_stream(L0["e2n"], var, L0["ewt"], f); out["ustream"] = f
    sf = np.zeros(var.shape[0]); o.calculate_dt(var, L0["vol"], sf); out["dt"] = sf.copy()
    out["min_dt"] = np.array([o.get_min_dt(sf)])
    o.compute_step_factor(var, L0["vol"], out["min_dt"][0], sf); out["step_factor"] = sf.copy()
    old = physical_state(var.shape[0], ff, seed + 2); out["in_old"] = old.copy()
    for rk in range(3):
        f = out["flux_edge"].copy(); v = var.copy()
        o.time_step(rk, sf, f, old, v)
        out[f"time_step_rk{rk}"] = v
        assert not f.any()
    res = np.zeros_like(var); o.residual(old, var, res); out["residual"] = res
    out["rms"] = np.array([o.calc_rms(res)])
    # restrict 0 -> 1 (coarse start state = far field so that childless nodes keep it, Q8)
    va = physical_state(L1["var"].shape[0], ff, seed + 3); out["in_var_above"] = va.copy()
    sc = np.zeros((va.shape[0], 2), dtype=np.int32)
    o.up_pre(L0["mg"], va, sc); o.up(L0["mg"], var, va, sc); o.up_post(va, sc)
    out["restrict"] = va.copy(); out["restrict_count"] = sc[:, 0].copy()
    # prolong 1 -> 0
    ra = np.random.default_rng(seed + 4).uniform(-1e-3, 1e-3, size=va.shape); out["in_res_above"] = ra.copy()
    v = var.copy(); o.down(L0["mg"], v, res, L0["coords"], ra, L1["coords"]); out["prolong"] = v
    return out


def cycle_vectors(o, mesh0, n_cycles):
    run = o.make_state(mesh0)
    run.init()
    rc, st = run.run(n_cycles)
    assert rc == 0
    out = {"min_dt": np.array([st.last_min_dt]), "rms": np.array([st.last_rms])}
    for l, a in enumerate(run.levels):
        out[f"var_L{l}"] = a["var"].copy()
        out[f"vol_L{l}"] = a["vol"].copy()
    return out


def main():
    if not orc.available("ref"):
        orc.build()
    if not orc.available("ref"):
        raise SystemExit("oracle/_ref is not built and /root/reference is absent: goldens can only be made in the build container")
    mg = _meshgen()
    o = orc.Oracle("ref")
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    for name, cycles in (("tiny", 3), ("small", 10)):
        mesh0 = [mg.zero_based(l) for l in mg.make_multigrid(name)["levels"]]
        np.savez_compressed(os.path.join(gold, f"{name}_cycles{cycles}.npz"), **cycle_vectors(o, mesh0, cycles))
    mesh0 = [mg.zero_based(l) for l in mg.make_multigrid("tiny")["levels"]]
    np.savez_compressed(os.path.join(gold, "tiny_loops.npz"), **loop_vectors(o, mesh0))
    np.savez_compressed(os.path.join(gold, "consts.npz"), consts=o.consts)
    print("wrote", sorted(os.listdir(gold)))


if __name__ == "__main__":
    main()
