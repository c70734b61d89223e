"""Index-set exactness (BASELINE.json north_star: "bit-exact edge colouring and partition/halo index sets").
The C++ planner is compared bit for bit with the independent restatement in oracle/plan_oracle.py, and the
colouring is checked for validity.  The planner is host code: these tests run on CPU through a planning-only
context (mgcfd_create(device=-1)), which can plan and answer plan queries but refuses every compute call."""
import numpy as np
import pytest

from conftest import mesh0



@pytest.fixture(scope="module")
def plan_oracle():
    import plan_oracle
    return plan_oracle


@pytest.mark.parametrize("name", ["tiny", "small", "medium"])
def test_renumbering_and_edge_order(pkg, meshgen, plan_oracle, name):
    mesh = meshgen.make_multigrid(name)
    lev0 = mesh0(meshgen, name)
    with pkg.MGCFD(mesh["levels"], init=False, device=-1) as gpu:
        for l, lev in enumerate(lev0):
            perm = gpu.plan_query(l, "node_perm")
            ref = plan_oracle.hilbert_renumber(lev["node_coordinates"])
            assert np.array_equal(perm, ref)
            assert np.array_equal(np.sort(perm), np.arange(perm.size))          # a permutation
            assert np.array_equal(gpu.plan_query(l, "edge_order"), plan_oracle.sort_edges(lev["edge-->node"], ref))


def test_renumbering_improves_locality(pkg, meshgen):
    mesh = meshgen.make_multigrid("medium")
    lev = mesh0(meshgen, "medium")[0]
    with pkg.MGCFD(mesh["levels"], init=False, device=-1) as gpu:
        perm = gpu.plan_query(0, "node_perm").astype(np.int64)
    e = lev["edge-->node"].astype(np.int64)
    before = np.abs(e[:, 0] - e[:, 1]).mean()
    after = np.abs(perm[e[:, 0]] - perm[e[:, 1]]).mean()
    assert after < before / 20


@pytest.mark.parametrize("name,bs", [("tiny", 256), ("small", 256), ("small", 64), ("medium", 256)])
def test_two_level_colouring_bit_exact(pkg, meshgen, plan_oracle, name, bs):
    mesh = meshgen.make_multigrid(name)
    lev0 = mesh0(meshgen, name)
    with pkg.MGCFD(mesh["levels"], init=False, device=-1, colour_block_edges=bs) as gpu:
        for l, lev in enumerate(lev0):
            perm, order = gpu.plan_query(l, "node_perm"), gpu.plan_query(l, "edge_order")
            tc, bc = gpu.plan_query(l, "edge_thread_colour"), gpu.plan_query(l, "edge_block_colour")
            rt, rb, nbc = plan_oracle.colour_edges(lev["edge-->node"], perm, order, bs)
            assert np.array_equal(tc, rt)
            assert np.array_equal(bc, rb)
            assert gpu.plan_query(l, "n_block_colours")[0] == nbc
            assert plan_oracle.check_colouring(lev["edge-->node"], order, bs, tc, bc)
            ex = gpu.plan_query(l, "colour_exec_edge")
            assert np.array_equal(np.sort(ex), np.arange(ex.size))              # every edge executes exactly once


@pytest.mark.parametrize("name,chunk", [("tiny", 256), ("small", 256), ("small", 64), ("medium", 128)])
def test_owner_chunks_bit_exact(pkg, meshgen, plan_oracle, name, chunk):
    mesh = meshgen.make_multigrid(name)
    lev0 = mesh0(meshgen, name)
    with pkg.MGCFD(mesh["levels"], init=False, device=-1, owner_chunk_nodes=chunk) as gpu:
        for l, lev in enumerate(lev0):
            perm = gpu.plan_query(l, "node_perm")
            starts = gpu.plan_query(l, "owner_chunk_start")
            hoff, hgid = gpu.plan_query(l, "owner_halo_off"), gpu.plan_query(l, "owner_halo_gid")
            eoff, efile = gpu.plan_query(l, "owner_edge_off"), gpu.plan_query(l, "owner_edge_file")
            max_loc, max_edges = chunk + (chunk * 3) // 2 + 64, chunk * 6
            rs, rh, re = plan_oracle.owner_chunks(lev["edge-->node"], perm, perm.size, chunk, max_loc, max_edges)
            assert list(starts) == rs
            for k in range(len(rs) - 1):
                assert list(hgid[hoff[k]:hoff[k + 1]]) == rh[k]
                assert list(efile[eoff[k]:eoff[k + 1]]) == re[k]
            # invariants: chunks tile the owned nodes; every edge appears in the chunk(s) of its endpoints only
            assert starts[0] == 0 and starts[-1] == perm.size and (np.diff(starts) > 0).all()
            counts = np.bincount(efile, minlength=lev["edge-->node"].shape[0])
            chunk_of = np.searchsorted(starts, perm[lev["edge-->node"]], side="right") - 1
            assert np.array_equal(counts, 1 + (chunk_of[:, 0] != chunk_of[:, 1]))
