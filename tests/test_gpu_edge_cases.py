"""Edge cases the index plumbing must survive: odd set sizes (16-byte alignment tails of the bulk copies), chunks
smaller than a warp, levels without boundary nodes, a level without edges, every boundary group, single-chunk decks."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
VARIANTS = ["owner", "emit", "gather", "colour", "atomic"]


def normwise(a, b):
    return np.abs(a - b).max(axis=0) / np.maximum(np.abs(b).max(axis=0), 1e-300)


def oracle_cycles(oracle_port, meshgen, mesh, cycles):
    lev0 = [meshgen.zero_based(l, base=mesh["base_array_index"]) for l in mesh["levels"]]
    run = oracle_port.make_state(lev0)
    run.init()
    assert run.run(cycles)[0] == 0
    return [a["var"] for a in run.levels]


@pytest.mark.parametrize("dims", [[(9, 7, 5, 900), (5, 5, 3, 190)], [(3, 3, 3, 60)], [(33, 3, 3, 700), (11, 3, 3, 260), (5, 3, 1, None)]])
@pytest.mark.parametrize("variant", VARIANTS)
def test_odd_sizes_and_tiny_decks(pkg, meshgen, oracle_port, dims, variant):
    mesh = meshgen.make_multigrid(("m6wing", dims, 17))
    ref = oracle_cycles(oracle_port, meshgen, mesh, 3)
    exact = variant in ("owner", "gather")
    for chunk in (128, 6):                         # 6: chunks far smaller than a warp, many of them
        with pkg.MGCFD(mesh["levels"], flux_variant=variant, exact_arith=exact, owner_chunk_nodes=chunk) as g:
            g.run_cycles(3)
            for l in range(len(ref)):
                got = g.fetch(l, "variables")
                if exact:
                    assert np.array_equal(got, ref[l]), (chunk, l)
                else:
                    assert (normwise(got, ref[l]) <= 1e-10).all(), (chunk, l)


def test_level_without_boundary_nodes_and_all_groups(pkg, meshgen, oracle_port):
    mesh = meshgen.make_multigrid(("m6wing", [(7, 6, 5, 600), (4, 4, 3, 120)], 23))
    lv1 = mesh["levels"][1]
    for k in ("bnd_node-->node", "bnd_node-->group"):
        lv1[k] = lv1[k][:0].copy()
    lv1["bnd_node_weights"] = lv1["bnd_node_weights"][:0].copy()
    g0 = mesh["levels"][0]["bnd_node-->group"]
    g0[:] = (np.arange(g0.shape[0]) % 13 - 2)[:, None]          # groups -2..10: wall, far field and no-op branches
    ref = oracle_cycles(oracle_port, meshgen, mesh, 2)
    with pkg.MGCFD(mesh["levels"], exact_arith=True) as g:
        g.run_cycles(2)
        for l in range(2):
            assert np.array_equal(g.fetch(l, "variables"), ref[l])


@pytest.mark.parametrize("variant", VARIANTS)
def test_level_without_edges_is_a_no_op_for_the_flux_loops(pkg, meshgen, variant):
    mesh = meshgen.make_multigrid(("m6wing", [(4, 3, 3, None)], 29))
    lev = mesh["levels"][0]
    lev["edge-->node"] = lev["edge-->node"][:0].copy()
    lev["edge_weights"] = lev["edge_weights"][:0].copy()
    with pkg.MGCFD(mesh["levels"], flux_variant=variant) as g:
        f = np.random.default_rng(1).uniform(-1, 1, size=(g.sizes[0][0], 5))
        g.set(0, "fluxes", f)
        g.compute_flux_edge(0)
        g.unstructured_stream(0)
        assert np.array_equal(g.fetch(0, "fluxes"), f)
        assert not g.fetch(0, "dummy_fluxes").any()


def test_fetch_set_round_trip_every_dat(pkg, meshgen):
    mesh = meshgen.make_multigrid(("m6wing", [(9, 7, 5, 900), (5, 5, 3, 190)], 3))
    rng = np.random.default_rng(5)
    with pkg.MGCFD(mesh["levels"]) as g:
        n = g.sizes[0][0]
        for name, shape in (("variables", (n, 5)), ("old_variables", (n, 5)), ("residuals", (n, 5)), ("fluxes", (n, 5)),
                            ("step_factors", (n,)), ("volumes", (n,))):
            a = rng.uniform(0.5, 2.0, size=shape)
            g.set(0, name, a)
            assert np.array_equal(g.fetch(0, name), a), name
        pinned = pkg.PinnedArray((n, 5))
        pinned.array[:] = rng.uniform(-1, 1, size=(n, 5))
        g.set(0, "variables", pinned.array)
        out = pkg.PinnedArray((n, 5))
        g.fetch_into(0, "variables", out.array)
        assert np.array_equal(out.array, pinned.array)
        pinned.free(); out.free()
        assert np.array_equal(g.fetch(0, "node_coordinates"), mesh["levels"][0]["node_coordinates"])
