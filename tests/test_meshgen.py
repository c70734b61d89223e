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
