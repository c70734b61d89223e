"""Partition / halo index sets (BASELINE.json north_star: bit-exact partition/halo index sets): the C++ partitioner
against the independent restatement in oracle/plan_oracle.py, plus the invariants of SURVEY.md 4.3-3.  Host code
only -- runs without a GPU."""
import numpy as np
import pytest

from conftest import mesh0


@pytest.fixture(scope="module")
def plan_oracle():
    import plan_oracle
    return plan_oracle


@pytest.mark.parametrize("name,n_ranks", [("tiny", 2), ("small", 2), ("small", 3), ("small", 4), ("small", 8), ("medium", 8)])
def test_partition_and_local_meshes_bit_exact(pkg, meshgen, plan_oracle, name, n_ranks):
    mesh = meshgen.make_multigrid(name)
    lev0 = mesh0(meshgen, name)
    parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], n_ranks)
    ref_parts = plan_oracle.partition_levels(lev0, n_ranks)
    for a, b in zip(parts, ref_parts):
        assert np.array_equal(a, b)
    sizes = np.bincount(parts[0], minlength=n_ranks)
    assert sizes.max() - sizes.min() <= n_ranks                 # bisection balances level 0
    locals_ = []
    for r in range(n_ranks):
        lm = pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, r, n_ranks)
        ref = plan_oracle.local_mesh(lev0, ref_parts, r)
        for l in range(len(lev0)):
            n_nodes, n_edges, n_bnd, n_owned = lm.sizes(l)
            gn = lm.query(l, "global_node")
            assert np.array_equal(gn, ref[l]["global_node"]) and n_owned == ref[l]["n_owned"]
            assert np.array_equal(lm.query(l, "global_edge"), ref[l]["global_edge"])
            assert np.array_equal(lm.query(l, "global_bnd"), ref[l]["global_bnd"])
            assert np.array_equal(lm.query(l, "edge_to_node").reshape(-1, 2), ref[l]["e2n"])
            if l + 1 < len(lev0):
                assert np.array_equal(lm.query(l, "node_to_mg_node"), ref[l]["mg"])
            nbr = lm.query(l, "neighbour_rank")
            assert list(nbr) == ref[l]["neighbour_rank"]
            ep, ei, ip = lm.query(l, "export_ptr"), lm.query(l, "export_idx"), lm.query(l, "import_ptr")
            for k, q in enumerate(nbr):
                assert list(gn[ei[ep[k]:ep[k + 1]]]) == ref[l]["exports"][q]
                assert list(gn[n_owned + ip[k]:n_owned + ip[k + 1]]) == ref[l]["imports"][q]
        locals_.append(lm)
    # invariants across ranks
    for l, lev in enumerate(lev0):
        e = lev["edge-->node"]
        seen_edges = np.zeros(e.shape[0], dtype=np.int64)
        owned_total = 0
        for r, lm in enumerate(locals_):
            gn, ge = lm.query(l, "global_node"), lm.query(l, "global_edge")
            n_owned = lm.sizes(l)[3]
            owned_total += n_owned
            seen_edges[ge] += 1
            assert np.isin(e[ge].ravel(), gn).all()             # owned + halo covers every referenced node
            nbr, ep, ei, ip = (lm.query(l, k) for k in ("neighbour_rank", "export_ptr", "export_idx", "import_ptr"))
            for k, q in enumerate(nbr):                         # my exports to q are exactly q's imports from me
                other = locals_[q]
                onbr, oip = other.query(l, "neighbour_rank"), other.query(l, "import_ptr")
                kk = list(onbr).index(r)
                ogn, ono = other.query(l, "global_node"), other.sizes(l)[3]
                assert np.array_equal(gn[ei[ep[k]:ep[k + 1]]], ogn[ono + oip[kk]:ono + oip[kk + 1]])
        assert owned_total == lev["node_coordinates"].shape[0]
        cut = parts[l][e[:, 0]] != parts[l][e[:, 1]]
        assert np.array_equal(seen_edges, 1 + cut)              # cut edges execute on both owners, others once


def test_single_rank_partition_is_the_whole_mesh(pkg, meshgen):
    mesh = meshgen.make_multigrid("tiny")
    parts = pkg.partition_levels(mesh["levels"], 1, 1)
    lm = pkg.LocalMesh(mesh["levels"], 1, parts, 0, 1)
    for l, lev in enumerate(mesh["levels"]):
        n = lev["node_coordinates"].shape[0]
        assert lm.sizes(l) == (n, lev["edge-->node"].shape[0], lev["bnd_node-->node"].shape[0], n)
        assert np.array_equal(lm.query(l, "global_node"), np.arange(n))
        assert lm.query(l, "neighbour_rank").size == 0


@pytest.mark.parametrize("name,n_ranks", [("tiny", 2), ("small", 2), ("small", 8), ("small", 5), ("small", 7), ("medium", 5)])
def test_kway_partition_bit_exact_and_better_than_geometric(pkg, meshgen, plan_oracle, name, n_ranks):
    """op_partition's k-way method (euler3d.cpp:340-375): recursive graph bisection with Fiduccia-Mattheyses refinement.
    Bit for bit against the restatement, balanced, and never a larger edge cut than the coordinate bisection it starts
    from: equal where the median falls between two grid planes of these structured-like decks (2, 8 ranks), 30-35 %
    smaller where a grid plane has to be shared (5, 7 ranks: the coordinate split scatters the shared plane)."""
    mesh = meshgen.make_multigrid(name)
    lev0 = mesh0(meshgen, name)
    parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], n_ranks, method="kway")
    ref_parts = plan_oracle.partition_levels(lev0, n_ranks, method="kway")
    for a, b in zip(parts, ref_parts):
        assert np.array_equal(a, b)
    geom = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], n_ranks)
    n = lev0[0]["node_coordinates"].shape[0]
    sizes = np.bincount(parts[0], minlength=n_ranks)
    assert sizes.min() > 0 and sizes.max() <= 1.03 * n / n_ranks + 2      # 0.5 % slack per bisection level
    cut_kway, cut_geom = plan_oracle.edge_cut(parts[0], lev0[0]["edge-->node"]), plan_oracle.edge_cut(geom[0], lev0[0]["edge-->node"])
    assert cut_kway <= cut_geom
    if n_ranks in (5, 7):
        assert cut_kway < 0.8 * cut_geom
    # the rank meshes built from a k-way partition satisfy the same invariants
    for r in range(n_ranks):
        lm = pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, r, n_ranks)
        ref = plan_oracle.local_mesh(lev0, ref_parts, r)
        for l in range(len(lev0)):
            assert np.array_equal(lm.query(l, "global_node"), ref[l]["global_node"])
            assert np.array_equal(lm.query(l, "global_edge"), ref[l]["global_edge"])


@pytest.mark.parametrize("method", ["block", "random", "inertial", "parmetis", "ptscotch", "geomkway"])
def test_partition_method_names(pkg, meshgen, method):
    """every library / method name of config.h:203-240 is accepted; each gives a balanced, complete partition"""
    mesh = meshgen.make_multigrid("small")
    parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], 4, method=method)
    n = mesh["levels"][0]["node_coordinates"].shape[0]
    sizes = np.bincount(parts[0], minlength=4)
    assert sizes.sum() == n and sizes.min() >= 0.95 * n / 4 - 2
    if method == "inertial":
        assert np.array_equal(parts[0], pkg.partition_levels(mesh["levels"], mesh["base_array_index"], 4)[0])
    if method in ("parmetis", "ptscotch", "geomkway"):
        assert np.array_equal(parts[0], pkg.partition_levels(mesh["levels"], mesh["base_array_index"], 4, method="kway")[0])
    if method == "block":
        assert (np.diff(parts[0]) >= 0).all()


def test_partition_unknown_method_is_an_error(pkg, meshgen):
    mesh = meshgen.make_multigrid("tiny")
    with pytest.raises(pkg.capi.MgcfdError):
        pkg.partition_levels(mesh["levels"], mesh["base_array_index"], 2, method="metis5")


@pytest.mark.parametrize("n_ranks", [2, 3, 8])
def test_export_tables_of_the_fused_push(pkg, meshgen, n_ranks):
    """the kernels of a multi-rank cycle push exported rows themselves: per owned node (internal
    numbering) the (destination slot, row) entries.  They must be exactly the export lists seen from the node's side --
    slot = position of the neighbour among those that receive rows, row = position in the export list for it, which
    is the row of the node in that neighbour's import range -- and both ends of every pair must agree on the counts.
    CPU only (planning-only contexts)."""
    mesh = meshgen.make_multigrid("small")
    parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], n_ranks)
    lms = [pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, r, n_ranks) for r in range(n_ranks)]
    ctxs = [pkg.MGCFD(local_mesh=lm, device=-1, init=False) for lm in lms]
    try:
        for l in range(len(mesh["levels"])):
            for r, (lm, g) in enumerate(zip(lms, ctxs)):
                perm = g.plan_query(l, "node_perm").astype(np.int64)           # internal index of local file node i
                nbr, ep, ei = lm.query(l, "neighbour_rank"), lm.query(l, "export_ptr"), lm.query(l, "export_idx")
                ptr = g.plan_query(l, "export_node_ptr")
                ent = g.plan_query(l, "export_node_ent").reshape(-1, 2)
                n_owned = g.n_owned[l]
                assert ptr.size == n_owned + 1 and ptr[0] == 0 and ptr[-1] == ent.shape[0] == ei.size
                expect = [[] for _ in range(n_owned)]
                slot = 0
                for k in range(nbr.size):
                    if ep[k + 1] == ep[k]:
                        continue
                    for j in range(ep[k], ep[k + 1]):
                        expect[perm[ei[j]]].append((slot, j - ep[k]))
                    slot += 1
                for v in range(n_owned):
                    assert [tuple(x) for x in ent[ptr[v]:ptr[v + 1]]] == expect[v], (l, r, v)
                # the receiving side expects as many rows from this rank as this rank sends
                for k in range(nbr.size):
                    q = int(nbr[k])
                    nbr_q, ip_q = lms[q].query(l, "neighbour_rank"), lms[q].query(l, "import_ptr")
                    kq = int(np.where(nbr_q == r)[0][0])
                    assert ip_q[kq + 1] - ip_q[kq] == ep[k + 1] - ep[k]
    finally:
        for g in ctxs:
            g.close()


@pytest.mark.parametrize("seed,n_ranks,method", [(11, 2, "geom"), (12, 3, "geom"), (13, 5, "kway"), (14, 8, "geom"), (15, 4, "random")])
def test_partition_of_irregular_two_level_decks_bit_exact(pkg, plan_oracle, seed, n_ranks, method):
    """decks that are not perturbed grids: random graphs with hubs, duplicated edges and isolated nodes, a random
    fine -> coarse map that leaves some coarse nodes childless (Q8) -- partition vectors, local meshes and halo lists of
    the C++ library against the restatement, plus the cross-rank invariants"""
    from test_plan_oracle import random_level
    rng = np.random.default_rng(seed)
    fine = random_level(rng, 400, 1800, hubs=2, dup=30, isolated=10)
    coarse = random_level(rng, 90, 300, hubs=0, dup=0, isolated=5)
    fine["node-->mg_node"] = rng.integers(1, 70 + 1, size=(400, 1)).astype(np.int32)       # coarse nodes 70..89 stay childless
    levels = [fine, coarse]
    lev0 = []
    for lv in levels:
        z = dict(lv)
        for k in ("edge-->node", "bnd_node-->node", "node-->mg_node"):
            if k in z:
                z[k] = z[k] - 1
        z["bnd_node-->node"] = z["bnd_node-->node"].reshape(-1)
        z["bnd_node-->group"] = z["bnd_node-->group"].reshape(-1)
        if "node-->mg_node" in z:
            z["node-->mg_node"] = z["node-->mg_node"].reshape(-1)
        lev0.append(z)
    parts = pkg.partition_levels(levels, 1, n_ranks, method=method)
    if method in ("geom", "kway"):
        ref_parts = plan_oracle.partition_levels(lev0, n_ranks, method=method)
    else:
        # the trivial partitioners are not restated: take their level-0 vector, restate the coarse rule and the halos on it
        ref_parts = [parts[0], plan_oracle.coarse_part(parts[0], lev0[0]["node-->mg_node"], 90, lev0[1]["edge-->node"],
                                                       np.asarray(lev0[1]["node_coordinates"]))]
    for a, b in zip(parts, ref_parts):
        assert np.array_equal(a, b)
        assert a.min() >= 0 and a.max() < n_ranks
    owned_total = [0, 0]
    for r in range(n_ranks):
        lm = pkg.LocalMesh(levels, 1, parts, r, n_ranks)
        ref = plan_oracle.local_mesh(lev0, ref_parts, r)
        for l in range(2):
            n_nodes, n_edges, n_bnd, n_owned = lm.sizes(l)
            gn = lm.query(l, "global_node")
            assert np.array_equal(gn, ref[l]["global_node"]) and n_owned == ref[l]["n_owned"]
            assert np.array_equal(lm.query(l, "global_edge"), ref[l]["global_edge"])
            assert np.array_equal(lm.query(l, "global_bnd"), ref[l]["global_bnd"])
            assert np.array_equal(lm.query(l, "edge_to_node").reshape(-1, 2), ref[l]["e2n"])
            if l == 0:
                assert np.array_equal(lm.query(l, "node_to_mg_node"), ref[l]["mg"])
            nbr = lm.query(l, "neighbour_rank")
            assert list(nbr) == ref[l]["neighbour_rank"]
            ep, ei, ip = lm.query(l, "export_ptr"), lm.query(l, "export_idx"), lm.query(l, "import_ptr")
            for k, q in enumerate(nbr):
                assert list(gn[ei[ep[k]:ep[k + 1]]]) == ref[l]["exports"][q]
                assert list(gn[n_owned + ip[k]:n_owned + ip[k + 1]]) == ref[l]["imports"][q]
            owned_total[l] += n_owned
            # the rank's context plans on it (chunks, export tables) without a GPU
        ctx = pkg.MGCFD(local_mesh=lm, device=-1, init=False)
        for l in range(2):
            assert int(ctx.plan_query(l, "owner_stats")[3]) <= 64
        ctx.close()
        lm.free()
    assert owned_total == [400, 90]
