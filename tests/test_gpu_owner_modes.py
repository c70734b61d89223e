"""Every execution mode of the owner flux kernel gives the reference's results: the default (one CTA per chunk,
bank-aware edge slots, node phase finishing from registers), the persistent variants with one and two
shared-memory stages, the staged epilogue and the plan-order slots.  Exact build: bit for bit against the
golden solutions of the reference's own arithmetic; fast build: the north_star tolerance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
STATE_TOL = 1e-10      # BASELINE.json north_star: per-variable flow state within 1e-10 relative

MODES = [
    {},                                                        # default
    {"MGCFD_OWNER_PIPE": "1"},                                 # persistent CTAs, one stage, next chunk prefetched into L2
    {"MGCFD_OWNER_PIPE": "2"},                                 # persistent CTAs, two shared-memory stages
    {"MGCFD_OWNER_EPILOGUE": "0", "MGCFD_OWNER_SLOT_SPLIT": "1"},   # node sums staged through shared memory
    {"MGCFD_OWNER_SLOTTING": "0"},                             # edge slots in plan (first-touch) order
    {"MGCFD_OWNER_PIPE": "2", "MGCFD_OWNER_THREADS": "256"},   # four threads per owned node in the fast build
]
KNOBS = ["MGCFD_OWNER_PIPE", "MGCFD_OWNER_EPILOGUE", "MGCFD_OWNER_SLOTTING", "MGCFD_OWNER_SLOT_SPLIT", "MGCFD_OWNER_THREADS"]


def normwise(a, b):
    return np.abs(a - b).max(axis=0) / np.maximum(np.abs(b).max(axis=0), 1e-300)


@pytest.mark.parametrize("mode", MODES, ids=lambda m: ",".join(f"{k[12:].lower()}={v}" for k, v in m.items()) or "default")
@pytest.mark.parametrize("exact", [True, False], ids=["exact", "fast"])
def test_owner_modes_match_golden(pkg, meshgen, golden, monkeypatch, mode, exact):
    for k in KNOBS:
        monkeypatch.delenv(k, raising=False)
    for k, v in mode.items():
        monkeypatch.setenv(k, v)
    g = golden("small_cycles10.npz")
    mesh = meshgen.make_multigrid("small")
    with pkg.MGCFD(mesh["levels"], flux_variant="owner", exact_arith=exact) as gpu:
        gpu.run_cycles(10)
        for l in range(len(mesh["levels"])):
            got, ref = gpu.fetch(l, "variables"), g[f"var_L{l}"]
            if exact:
                assert np.array_equal(got, ref), (mode, l, np.abs(got - ref).max())
            else:
                assert (normwise(got, ref) <= STATE_TOL).all(), (mode, l, normwise(got, ref))
                assert gpu.validate(l, ref) == 0


@pytest.mark.parametrize("mode", MODES[:3], ids=["default", "pipe1", "pipe2"])
def test_owner_modes_unfused_flux_loop(pkg, meshgen, monkeypatch, mode):
    """the stand-alone compute_flux_edge call site (accumulating and overwriting launches) in every kernel mode:
    bit-identical flux arrays in the exact build"""
    mesh = meshgen.make_multigrid("small")
    out = []
    for m in ({"MGCFD_OWNER_PIPE": "0", "MGCFD_OWNER_SLOTTING": "0"}, mode):
        for k in KNOBS:
            monkeypatch.delenv(k, raising=False)
        for k, v in m.items():
            monkeypatch.setenv(k, v)
        with pkg.MGCFD(mesh["levels"], flux_variant="owner", exact_arith=True) as gpu:
            gpu.compute_flux_edge(0)          # flux known to be zero: overwriting launch
            gpu.compute_flux_edge(0)          # accumulating launch
            out.append(gpu.fetch(0, "fluxes"))
    assert np.abs(out[0]).max() > 0
    assert np.array_equal(out[0], out[1])
