"""The synthetic decks have the sizes BASELINE.json / SURVEY.md 8d name and the reference's file layout."""
import numpy as np


def test_m6_sizes(meshgen):
    m = meshgen.make_multigrid("m6")
    nodes = [l["node_coordinates"].shape[0] for l in m["levels"]]
    edges = [l["edge-->node"].shape[0] for l in m["levels"]]
    assert nodes == [300_000, 165_000, 111_000, 81_000]          # README.md:97
    assert edges[0] == 930_000 and sum(edges) == 2_438_117       # analyse-output-data.py:116
    for l in m["levels"][:-1]:
        assert l["node-->mg_node"].shape == (l["node_coordinates"].shape[0], 1)
    assert "node-->mg_node" not in m["levels"][-1]


def test_layout_and_determinism(meshgen):
    a, b = meshgen.make_multigrid("small"), meshgen.make_multigrid("small")
    for la, lb in zip(a["levels"], b["levels"]):
        for k in la:
            assert np.array_equal(la[k], lb[k])
    l0 = a["levels"][0]
    n = l0["node_coordinates"].shape[0]
    assert l0["edge-->node"].dtype == np.int32 and l0["edge-->node"].min() == 1 and l0["edge-->node"].max() == n   # Q12
    assert l0["bnd_node-->group"].dtype == np.int32
    assert set(np.unique(l0["bnd_node-->group"])) == {0, 1, 2, 3, 5, 9}      # every dispatch branch of flux.h:29-37
    e = l0["edge-->node"]
    assert (e[:, 0] != e[:, 1]).all()
    key = np.sort(e, axis=1)
    assert np.unique(key, axis=0).shape[0] == e.shape[0]                      # no duplicate edges
    cn = a["levels"][1]["node_coordinates"].shape[0]
    mg = l0["node-->mg_node"][:, 0] - 1
    assert mg.min() >= 0 and mg.max() < cn
    assert np.setdiff1d(np.arange(cn), mg).size > 0                           # childless coarse nodes exist (Q8)
    assert not np.array_equal(np.sort(e[:, 0]), e[:, 0])                      # file order is shuffled


import pytest


@pytest.mark.parametrize("n_ranks", [2, 3, 4, 5])
def test_slab_deck_generated_per_rank_equals_the_partitioned_whole(pkg, meshgen, n_ranks):
    """BASELINE.json configs[4] ("generated per-partition, never materialised on one host"): a rank's
    slab generated alone is exactly what mgcfd_local_mesh_build cuts out of the whole deck -- nodes, edges, weights,
    boundary entries, halo / export / import lists, bit for bit (12 x-planes over 2..5 ranks, divisible or not)"""
    mesh = meshgen.make_slab_global("slab_test")
    assert tuple(mesh["levels"][0][k].shape[0] for k in ("node_coordinates", "edge-->node", "bnd_node-->node")) == meshgen.slab_sizes("slab_test")
    part = meshgen.slab_part("slab_test", n_ranks)
    owned_total = 0
    for r in range(n_ranks):
        lm = pkg.LocalMesh(mesh["levels"], 1, [part], r, n_ranks)
        d = meshgen.make_slab_rank("slab_test", r, n_ranks)
        rm = pkg.RankMesh(d)
        assert lm.sizes(0) == rm.sizes(0)
        for what in ("global_node", "neighbour_rank", "export_ptr", "export_idx", "import_ptr", "edge_to_node"):
            assert np.array_equal(lm.query(0, what), rm.query(0, what)), (r, what)
        assert np.array_equal(lm.query(0, "global_edge"), d["global_edge"]) and np.array_equal(lm.query(0, "global_bnd"), d["global_bnd"])
        v = lm.level(0).contents
        for ptr, n, ref in ((v.node_coordinates, v.n_nodes * 3, d["node_coordinates"]), (v.edge_weights, v.n_edges * 3, d["edge_weights"]),
                            (v.bnd_node_weights, v.n_bnd_nodes * 3, d["bnd_node_weights"]),
                            (v.bnd_node_to_group, v.n_bnd_nodes, d["bnd_node-->group"]), (v.bnd_node_to_node, v.n_bnd_nodes, d["bnd_node-->node"])):
            assert np.array_equal(np.ctypeslib.as_array(ptr, shape=(n,)), ref.ravel())
        owned_total += d["n_owned"]
    assert owned_total == meshgen.slab_sizes("slab_test")[0]


def test_slab_deck_shape_of_the_150m_config(meshgen):
    n, e, b = meshgen.slab_sizes("rotor37_150m")
    assert n == 150_000_000 and e == 449_150_000 and b == 2 * (500 * 500 + 600 * 500 * 2)      # "~449 M axis edges"
    x = [meshgen.slab_owner_planes(600, r, 8) for r in range(8)]
    assert x[0] == (0, 75) and x[7] == (525, 600) and all(hi - lo == 75 for lo, hi in x)
    # the hashed quantities are reproducible and well spread
    u = meshgen._hash_u01(np.arange(100000), 2, 100)
    assert 0.49 < u.mean() < 0.51 and u.min() >= 0 and u.max() < 1 and np.array_equal(u, meshgen._hash_u01(np.arange(100000), 2, 100))
