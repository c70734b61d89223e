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


@pytest.mark.parametrize("n_ranks", [2, 4])
def test_p2p_unfused_exchange_still_bit_identical(pkg, meshgen, golden, n_ranks, monkeypatch):
    """MGCFD_FUSED_PUSH=0: the round-1 schedule (pack + signal kernels on the communication stream, stages split into
    export chunks and interior chunks) stays available and exact"""
    monkeypatch.setenv("MGCFD_FUSED_PUSH", "0")
    mesh = meshgen.make_multigrid("small")
    got, _ = run_decomposed(pkg, mesh, n_ranks, 10, exact_arith=True, p2p=True)
    g = golden("small_cycles10.npz")
    for l in range(len(mesh["levels"])):
        assert np.array_equal(got[l], g[f"var_L{l}"])


def test_fused_push_needs_fewer_launches(pkg, meshgen, monkeypatch):
    """the stage kernels push their exported rows and hand-shake themselves: no pack / signal launches per stage and no
    split into export and interior launches"""
    mesh = meshgen.make_multigrid("small")
    counts = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("MGCFD_FUSED_PUSH", fused)
        parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], 2)
        lms = [pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, r, 2) for r in range(2)]
        ranks = [pkg.MGCFD(local_mesh=lm, device=0) for lm in lms]
        try:
            pkg.group_enable_p2p(ranks)
            pkg.group_run_cycles(ranks, 1)
            before = ranks[0].kernel_launches()
            pkg.group_run_cycles(ranks, 2)
            counts[fused] = (ranks[0].kernel_launches() - before) / 2
        finally:
            for g in ranks:
                g.close()
    n_levels = len(mesh["levels"])
    visits = 2 * n_levels - 2
    assert counts["1"] <= counts["0"] - 3 * 3 * visits + visits, counts      # 3 launches fewer per stage, one wait per visit


@pytest.mark.parametrize("p2p", [False, True], ids=["events", "p2p"])
def test_bad_values_on_one_rank_fail_the_group_run(pkg, meshgen, p2p):
    """euler3d.cpp:544-548 on a decomposed run: NaNs that appear on one rank make the run return MGCFD_ERR_BAD_VALS
    (round 1 returned OK from multi-rank runs without looking at the flags)"""
    mesh = meshgen.make_multigrid("small")
    parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], 2)
    lms = [pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, r, 2) for r in range(2)]
    ranks = [pkg.MGCFD(local_mesh=lm, device=0) for lm in lms]
    try:
        if p2p:
            pkg.group_enable_p2p(ranks)
        v = ranks[1].fetch(0, "variables")
        v[ranks[1].n_owned[0] // 2, 0] = np.nan
        ranks[1].set(0, "variables", v)
        with pytest.raises(pkg.MgcfdError) as ei:
            pkg.group_run_cycles(ranks, 1)
        assert ei.value.code in (-4, -5)
    finally:
        for g in ranks:
            g.close()
