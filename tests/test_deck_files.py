"""Deck files: the container that stands in for the reference's HDF5 level files (same dataset names / shapes /
dtypes, euler3d.cpp:248-312) and the input.dat deck (io.h:28-205)."""
import os

import numpy as np


def test_container_round_trip(tmp_path, meshgen):
    mesh = meshgen.make_multigrid("tiny")
    deck = meshgen.write_deck(str(tmp_path), mesh)
    text = open(deck).read()
    assert "num_levels = 3" in text and "mesh_name = m6wing" in text and "base_array_index = 1" in text and "[levels]" in text
    for l, lev in enumerate(mesh["levels"]):
        back = meshgen.read_container(os.path.join(str(tmp_path), f"mesh.L{l}.mgb"))
        assert set(back) == set(lev)
        for k in lev:
            assert back[k].dtype == lev[k].dtype and np.array_equal(back[k], lev[k])
    p = meshgen.write_solution(str(tmp_path), 1, 10, np.arange(20.0).reshape(4, 5))
    assert os.path.basename(p) == "solution.variables.L1.cycles=10.mgb"              # Q14 naming
    assert np.array_equal(meshgen.read_container(p)["p_variables_result_L1"], np.arange(20.0).reshape(4, 5))
