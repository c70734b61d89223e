"""Round-2 preparation (branch r2-prep): the lean owner kernel (MGCFD_OWNER_LEAN=1).  NOT yet run on a GPU; enable with
MGCFD_TEST_EXPERIMENTAL=1."""
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("MGCFD_TEST_EXPERIMENTAL") != "1", reason="experimental kernel")]


def test_lean_kernel_matches_golden(pkg, meshgen, golden, monkeypatch):
    monkeypatch.setenv("MGCFD_OWNER_LEAN", "1")
    g = golden("small_cycles10.npz")
    mesh = meshgen.make_multigrid("small")
    with pkg.MGCFD(mesh["levels"], flux_variant="owner") as gpu:
        gpu.run_cycles(10)
        for l in range(len(mesh["levels"])):
            got, ref = gpu.fetch(l, "variables"), g[f"var_L{l}"]
            assert (np.abs(got - ref).max(axis=0) <= 1e-10 * np.abs(ref).max(axis=0)).all()
            assert gpu.validate(l, ref) == 0


def test_per_rank_slab_decks_run_like_the_whole_deck(pkg, meshgen):
    """round-2 preparation: contexts built from meshgen.make_slab_rank / RankMesh (BASELINE configs[4] path) driven as
    virtual ranks reproduce the undecomposed run bit for bit in the exact build"""
    mesh = meshgen.make_slab_global("slab_test")
    with pkg.MGCFD(mesh["levels"], exact_arith=True) as single:
        single.run_cycles(3)
        ref = single.fetch(0, "variables")
    for n_ranks in (2, 3):
        rms = [pkg.RankMesh(meshgen.make_slab_rank("slab_test", r, n_ranks)) for r in range(n_ranks)]
        ranks = [pkg.MGCFD(local_mesh=rm, device=0, exact_arith=True) for rm in rms]
        try:
            pkg.group_run_cycles(ranks, 3)
            full = np.full_like(ref, np.nan)
            for r, g in enumerate(ranks):
                gn = rms[r].query(0, "global_node")
                no = g.n_owned[0]
                full[gn[:no]] = g.fetch(0, "variables")[:no]
            assert np.array_equal(full, ref)
        finally:
            for g in ranks:
                g.close()
