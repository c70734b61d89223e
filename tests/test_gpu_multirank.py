"""Multi-GPU = single-GPU parity.  The ranks are contexts of one process driven in lock step by
mgcfd_group_run_cycles; on a one-GPU box they all live on device 0 ("N virtual ranks"), which exercises the same
partitioner, halo lists, pack kernels, peer copies and min_dt reduction as a run over N devices.

In the exact build the decomposed run must reproduce the undecomposed one BIT FOR BIT: cut edges are recomputed on
both owners, every owned node sums its increments in ascending file order of the undecomposed mesh, and restrict
sums children in that order too."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def normwise(a, b):
    return np.abs(a - b).max(axis=0) / np.maximum(np.abs(b).max(axis=0), 1e-300)


def run_decomposed(pkg, mesh, n_ranks, cycles, devices=None, p2p=False, **kw):
    parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], n_ranks)
    lms = [pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, r, n_ranks) for r in range(n_ranks)]
    ranks = [pkg.MGCFD(local_mesh=lm, device=(devices[r] if devices else 0), **kw) for r, lm in enumerate(lms)]
    try:
        if p2p and n_ranks > 1:
            pkg.group_enable_p2p(ranks)
        pkg.group_run_cycles(ranks, cycles)
        out = []
        for l, lev in enumerate(mesh["levels"]):
            full = np.full((lev["node_coordinates"].shape[0], 5), np.nan)
            for r, g in enumerate(ranks):
                gn = lms[r].query(l, "global_node")
                no = g.n_owned[l]
                full[gn[:no]] = g.fetch(l, "variables")[:no]
            out.append(full)
        halo = sum(g.halo_bytes_sent() for g in ranks)
        return out, halo
    finally:
        for g in ranks:
            g.close()


@pytest.mark.parametrize("p2p", [False, True], ids=["events", "p2p"])
@pytest.mark.parametrize("n_ranks", [2, 4, 8])
def test_virtual_ranks_bit_identical_to_single(pkg, meshgen, golden, n_ranks, p2p):
    """p2p: the direct peer-store transport (pack kernel writes into the neighbours' halo ranges, epoch flags, min_dt
    mailboxes) instead of events + peer copies"""
    mesh = meshgen.make_multigrid("small")
    got, halo = run_decomposed(pkg, mesh, n_ranks, 10, exact_arith=True, p2p=p2p)
    g = golden("small_cycles10.npz")
    assert halo > 0
    for l in range(len(mesh["levels"])):
        assert not np.isnan(got[l]).any()                       # every node is owned by exactly one rank
        assert np.array_equal(got[l], g[f"var_L{l}"]), (l, np.abs(got[l] - g[f"var_L{l}"]).max())


@pytest.mark.parametrize("p2p", [False, True], ids=["events", "p2p"])
@pytest.mark.parametrize("n_ranks", [2, 3, 8])
def test_virtual_ranks_fast_build(pkg, meshgen, n_ranks, p2p):
    mesh = meshgen.make_multigrid("medium")
    with pkg.MGCFD(mesh["levels"]) as single:
        single.run_cycles(3)
        ref = [single.fetch(l, "variables") for l in range(len(mesh["levels"]))]
    got, _ = run_decomposed(pkg, mesh, n_ranks, 3, p2p=p2p)
    for l in range(len(ref)):
        assert (normwise(got[l], ref[l]) <= 1e-13).all(), (l, normwise(got[l], ref[l]))


def test_single_rank_group_is_plain_run(pkg, meshgen):
    mesh = meshgen.make_multigrid("tiny")
    got, halo = run_decomposed(pkg, mesh, 1, 3, exact_arith=True)
    with pkg.MGCFD(mesh["levels"], exact_arith=True) as single:
        single.run_cycles(3)
        for l in range(len(mesh["levels"])):
            assert np.array_equal(got[l], single.fetch(l, "variables"))
    assert halo == 0


def test_partitioned_context_refuses_plain_run(pkg, meshgen):
    mesh = meshgen.make_multigrid("tiny")
    parts = pkg.partition_levels(mesh["levels"], 1, 2)
    lm = pkg.LocalMesh(mesh["levels"], 1, parts, 0, 2)
    with pkg.MGCFD(local_mesh=lm) as g:
        with pytest.raises(pkg.MgcfdError):
            g.run_cycles(1)          # no communicator: a partition cannot advance alone
