"""Programmatic dependent launch of the cycle's kernels (MGCFD_PDL, csrc/internal.h): visit_begin, the fused
Runge-Kutta stage, restrict and prolong become resident while their predecessor drains and block in
griddepcontrol.wait before their first access to mutable data.  The arithmetic is untouched, so every result must be
BIT-identical with the knob on and off -- in CUDA-graph replay, in plain stream launches, with -b, through the
host-buffer call, and decomposed over virtual ranks with the fused halo push."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def run_single(pkg, mesh, cycles, **kw):
    with pkg.MGCFD(mesh["levels"], **kw) as gpu:
        for n in cycles:
            gpu.run_cycles(n)
        return [gpu.fetch(l, "variables") for l in range(len(mesh["levels"]))], gpu.kernel_launches()


@pytest.mark.parametrize("graphs", [True, False], ids=["graph", "stream"])
@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_pdl_bit_identical_single(pkg, meshgen, golden, monkeypatch, graphs, exact):
    mesh = meshgen.make_multigrid("small")
    out = {}
    for on in ("0", "1"):
        monkeypatch.setenv("MGCFD_PDL", on)
        out[on], launches = run_single(pkg, mesh, (1, 2, 7), exact_arith=exact, graphs=graphs)     # odd + even pieces: both parity graphs
    g = golden("small_cycles10.npz")
    for l in range(len(mesh["levels"])):
        assert np.array_equal(out["0"][l], out["1"][l]), (l, np.abs(out["0"][l] - out["1"][l]).max())
        if exact:
            assert np.array_equal(out["1"][l], g[f"var_L{l}"])


def test_pdl_bit_identical_larger_levels(pkg, meshgen, monkeypatch):
    """levels of several waves of CTAs (215K nodes): the dependents really start under a draining predecessor"""
    mesh = meshgen.make_multigrid("medium")
    out = {}
    for on in ("0", "1"):
        monkeypatch.setenv("MGCFD_PDL", on)
        out[on], _ = run_single(pkg, mesh, (5,))
    for l in range(len(mesh["levels"])):
        assert np.isfinite(out["1"][l]).all()
        assert np.array_equal(out["0"][l], out["1"][l])


def test_pdl_with_mem_bound_kernel_and_host_buffers(pkg, meshgen, monkeypatch):
    """-b puts a kernel that knows nothing of the protocol between the stages; mgcfd_run_cycles_host adds copies and
    events on a second stream around the level visits"""
    mesh = meshgen.make_multigrid("small")
    nl = len(mesh["levels"])
    out = {}
    for on in ("0", "1"):
        monkeypatch.setenv("MGCFD_PDL", on)
        res, _ = run_single(pkg, mesh, (3,), measure_mem_bound=True)
        with pkg.MGCFD(mesh["levels"]) as gpu:
            pin_in = [pkg.PinnedArray((s[0], 5)) for s in gpu.sizes]
            pin_out = [pkg.PinnedArray((s[0], 5)) for s in gpu.sizes]
            for l in range(nl):
                pin_in[l].array[:] = gpu.fetch(l, "variables")
            gpu.run_cycles_host(3, [p.array for p in pin_in], [p.array for p in pin_out])       # page-locked: the pipelined path
            got = [p.array.copy() for p in pin_out]
            for p in pin_in + pin_out:
                p.free()
        out[on] = (res, got)
    for l in range(nl):
        assert np.array_equal(out["0"][0][l], out["1"][0][l])
        assert np.array_equal(out["0"][1][l], out["1"][1][l])
        assert np.array_equal(out["1"][0][l], out["1"][1][l])      # -b changes nothing; host-buffer call = resident call


@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
@pytest.mark.parametrize("n_ranks", [2, 4])
def test_pdl_virtual_ranks_fused_push(pkg, meshgen, golden, monkeypatch, n_ranks, exact):
    from test_gpu_multirank import run_decomposed
    mesh = meshgen.make_multigrid("small")
    out = {}
    for on in ("0", "1"):
        monkeypatch.setenv("MGCFD_PDL", on)
        out[on], halo = run_decomposed(pkg, mesh, n_ranks, 10, exact_arith=exact, p2p=True)
        assert halo > 0
    g = golden("small_cycles10.npz")
    for l in range(len(mesh["levels"])):
        assert not np.isnan(out["1"][l]).any()
        assert np.array_equal(out["0"][l], out["1"][l])
        if exact:
            assert np.array_equal(out["1"][l], g[f"var_L{l}"])
