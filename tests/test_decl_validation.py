"""Argument checking at the drop-in boundary (include/mgcfd_b200.h) on a planning-only context (no GPU needed): OP2's
op_decl_set / op_decl_map / op_decl_dat abort on sizes, null pointers and map entries that do not fit
(euler3d.cpp:248-312 goes through them); here every such declaration returns MGCFD_ERR_ARG with a message and leaves
the process alive."""
import ctypes as C

import numpy as np
import pytest


def make_ctx(pkg, n_levels=2, **opt_kw):
    lib = pkg.load_library()
    opt = pkg.capi.Options()
    lib.mgcfd_default_options(C.byref(opt))
    for k, v in opt_kw.items():
        setattr(opt, k, v)
    ctx = C.c_void_p()
    assert lib.mgcfd_create(C.byref(ctx), -1, n_levels, C.byref(opt)) == 0
    return lib, ctx


@pytest.fixture()
def tiny(pkg):
    return pkg.meshgen.make_multigrid("tiny")


def err(lib, ctx):
    return lib.mgcfd_last_error(ctx).decode()


def test_well_formed_level_is_accepted(pkg, tiny):
    lib, ctx = make_ctx(pkg, len(tiny["levels"]))
    keep = []
    for l, lev in enumerate(tiny["levels"]):
        assert lib.mgcfd_decl_level(ctx, l, C.byref(pkg.capi._level_struct(lev, keep)), tiny["base_array_index"]) == 0
    lib.mgcfd_destroy(ctx)


@pytest.mark.parametrize("field", ["node_coordinates", "edge_to_node", "edge_weights", "bnd_node_to_node",
                                   "bnd_node_to_group", "bnd_node_weights"])
def test_null_dataset_is_an_error_not_a_crash(pkg, tiny, field):
    lib, ctx = make_ctx(pkg, len(tiny["levels"]))
    keep = []
    h = pkg.capi._level_struct(tiny["levels"][0], keep)
    setattr(h, field, type(getattr(h, field))())
    assert lib.mgcfd_decl_level(ctx, 0, C.byref(h), tiny["base_array_index"]) == -1
    assert "null" in err(lib, ctx)
    lib.mgcfd_destroy(ctx)


def test_sizes_levels_and_call_order(pkg, tiny):
    lib, ctx = make_ctx(pkg, len(tiny["levels"]))
    keep = []
    h = pkg.capi._level_struct(tiny["levels"][0], keep)
    assert lib.mgcfd_decl_level(ctx, -1, C.byref(h), 1) == -1
    assert lib.mgcfd_decl_level(ctx, len(tiny["levels"]), C.byref(h), 1) == -1
    assert lib.mgcfd_decl_level(ctx, 0, None, 1) == -1
    for fld in ("n_nodes", "n_edges", "n_bnd_nodes"):
        g = pkg.capi._level_struct(tiny["levels"][0], keep)
        setattr(g, fld, -1)
        assert lib.mgcfd_decl_level(ctx, 0, C.byref(g), 1) == -1 and "negative" in err(lib, ctx)
    g = pkg.capi._level_struct(tiny["levels"][0], keep)
    g.n_owned_nodes = g.n_nodes + 1
    assert lib.mgcfd_decl_level(ctx, 0, C.byref(g), 1) == -1
    g = pkg.capi._level_struct(tiny["levels"][0], keep)
    g.n_owned_nodes = g.n_nodes - 1                    # halo nodes need import lists
    assert lib.mgcfd_decl_level(ctx, 0, C.byref(g), 1) == -1 and "halo" in err(lib, ctx)
    # planning needs the constants first, and every level
    assert lib.mgcfd_plan(ctx) == -1
    lib.mgcfd_destroy(ctx)


def test_map_entries_are_range_checked(pkg, tiny):
    lev0 = tiny["levels"][0]
    n = lev0["node_coordinates"].shape[0]
    base = tiny["base_array_index"]
    cases = []
    bad = dict(lev0); e = lev0["edge-->node"].copy(); e[3, 1] = n + base; bad["edge-->node"] = e
    cases.append((bad, base, "edge-->node"))
    cases.append((dict(lev0), base + 1, "edge-->node"))                  # wrong base_array_index: an entry becomes -1
    bad = dict(lev0); e = lev0["edge-->node"].copy(); e[5, 1] = e[5, 0]; bad["edge-->node"] = e
    cases.append((bad, base, "self edge"))
    bad = dict(lev0); b = lev0["bnd_node-->node"].copy(); b[0] = n + base; bad["bnd_node-->node"] = b
    cases.append((bad, base, "bnd_node-->node"))
    for lev, b, what in cases:
        lib, ctx = make_ctx(pkg, len(tiny["levels"]))
        keep = []
        assert lib.mgcfd_decl_level(ctx, 0, C.byref(pkg.capi._level_struct(lev, keep)), b) == -1
        assert what in err(lib, ctx), err(lib, ctx)
        lib.mgcfd_destroy(ctx)


def test_mg_map_is_checked_at_plan_time(pkg, tiny):
    """node-->mg_node refers to the NEXT level's nodes, so its range can only be checked once every level is declared"""
    levels = [dict(l) for l in tiny["levels"]]
    mg = levels[0]["node-->mg_node"].copy()
    mg[7] = levels[1]["node_coordinates"].shape[0] + tiny["base_array_index"]
    levels[0]["node-->mg_node"] = mg
    with pytest.raises(pkg.MgcfdError) as ei:
        pkg.MGCFD(levels, base_array_index=tiny["base_array_index"], device=-1, init=False)
    assert ei.value.code == -1 and "mg_node" in str(ei.value)
    # a map on the coarsest level has nothing to point at
    levels = [dict(l) for l in tiny["levels"]]
    levels[-1]["node-->mg_node"] = np.ones((levels[-1]["node_coordinates"].shape[0], 1), dtype=np.int32)
    with pytest.raises(pkg.MgcfdError) as ei:
        pkg.MGCFD(levels, base_array_index=tiny["base_array_index"], device=-1, init=False)
    assert ei.value.code == -1
    # and a missing map on a finer level is caught by the planner
    levels = [dict(l) for l in tiny["levels"]]
    del levels[0]["node-->mg_node"]
    with pytest.raises(pkg.MgcfdError) as ei:
        pkg.MGCFD(levels, base_array_index=tiny["base_array_index"], device=-1, init=False)
    assert ei.value.code == -1 and "mg_node" in str(ei.value)


def test_halo_lists_are_checked(pkg, tiny):
    parts = pkg.partition_levels(tiny["levels"], tiny["base_array_index"], 2)
    lm = pkg.LocalMesh(tiny["levels"], tiny["base_array_index"], parts, 0, 2)
    # a partitioned level on a context that was created for one rank
    lib, ctx = make_ctx(pkg, lm.n_levels)
    assert lib.mgcfd_decl_level(ctx, 0, lm.level(0), 0) == -1 and "n_ranks" in err(lib, ctx)
    lib.mgcfd_destroy(ctx)
    # neighbour rank out of range for the context's world
    lib, ctx = make_ctx(pkg, lm.n_levels, rank=0, n_ranks=2)
    assert lib.mgcfd_decl_level(ctx, 0, lm.level(0), 0) == 0
    lib.mgcfd_destroy(ctx)
    lib, ctx = make_ctx(pkg, lm.n_levels, rank=1, n_ranks=2)       # rank 0's lists name rank 1 as the neighbour: not me
    assert lib.mgcfd_decl_level(ctx, 0, lm.level(0), 0) == -1 and "neighbour" in err(lib, ctx)
    lib.mgcfd_destroy(ctx)
    lm.free()


def test_null_context_is_an_error_on_every_entry_point(pkg):
    """every export whose first parameter is the context: NULL gives an error code (or 0 / NULL for the getters), never a
    crash -- the reference's op_* calls abort on a NULL set / dat handle, a library must not take the process down"""
    import os
    import re
    from conftest import ROOT
    hdr = open(os.path.join(ROOT, "include", "mgcfd_b200.h")).read()
    decls = re.findall(r"^([a-z][a-z \*]*?)\b(mgcfd_[a-z0-9_]+)\((?:const )?mgcfd_ctx \*ctx[,)]", hdr, flags=re.M)
    names = {n: r.strip() for r, n in decls}
    assert len(names) >= 40, sorted(names)
    lib = pkg.load_library()
    import subprocess, sys, json
    # in a child process: a crash there is a failed test here, not a dead test session
    code = r'''
import ctypes as C, json, sys
sys.path.insert(0, %r)
import __graft_entry__ as ge
pkg = ge.load_package()
lib = pkg.load_library()
names = json.loads(%r)
out = {}
for n, ret in names.items():
    f = getattr(lib, n)
    f.restype = C.c_void_p if "*" in ret else (C.c_longlong if "long long" in ret else (None if ret == "void" else C.c_int))
    f.argtypes = [C.c_void_p] * 8
    r = f(None, None, None, None, None, None, None, None)
    out[n] = r
print(json.dumps(out))
''' % (ROOT, json.dumps(names))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, (p.returncode, p.stderr[-1500:])
    res = json.loads(p.stdout.strip().splitlines()[-1])
    for n, ret in names.items():
        if ret == "int":
            assert res[n] is not None and res[n] < 0, (n, res[n])
        elif ret == "long long":
            assert res[n] <= 0, (n, res[n])
        elif "*" in ret and n != "mgcfd_last_error":
            assert not res[n], (n, res[n])


def test_local_mesh_build_refuses_what_does_not_fit(pkg, tiny):
    """mgcfd_local_mesh_build indexes the deck by the partition vectors and the maps: every one of them is checked first"""
    levels, base = tiny["levels"], tiny["base_array_index"]
    parts = pkg.partition_levels(levels, base, 2)
    pkg.LocalMesh(levels, base, parts, 1, 2).free()                        # the well-formed case
    n0 = levels[0]["node_coordinates"].shape[0]

    def bad_build(lv, pt, rank=0, n_ranks=2):
        with pytest.raises(pkg.MgcfdError) as ei:
            pkg.LocalMesh(lv, base, pt, rank, n_ranks)
        assert ei.value.code == -1

    p = [q.copy() for q in parts]; p[0][5] = 2                             # owner rank outside the world
    bad_build(levels, p)
    p = [q.copy() for q in parts]; p[1][0] = -1
    bad_build(levels, p)
    bad_build(levels, parts, rank=2)                                       # this rank outside the world
    lv = [dict(l) for l in levels]; e = lv[0]["edge-->node"].copy(); e[0, 0] = n0 + base; lv[0]["edge-->node"] = e
    bad_build(lv, parts)
    lv = [dict(l) for l in levels]; b = lv[0]["bnd_node-->node"].copy(); b[0] = base - 1; lv[0]["bnd_node-->node"] = b
    bad_build(lv, parts)
    lv = [dict(l) for l in levels]; m = lv[0]["node-->mg_node"].copy()
    m[3] = levels[1]["node_coordinates"].shape[0] + base; lv[0]["node-->mg_node"] = m
    bad_build(lv, parts)
    lv = [dict(l) for l in levels]; del lv[0]["node-->mg_node"]           # a finer level without its map
    bad_build(lv, parts)


def test_partitioners_check_their_arguments(pkg, tiny):
    import ctypes as C
    lib = pkg.load_library()
    lev = tiny["levels"][0]
    xyz = np.ascontiguousarray(lev["node_coordinates"], dtype=np.float64)
    e2n = np.ascontiguousarray(lev["edge-->node"], dtype=np.int32)
    n, E = xyz.shape[0], e2n.shape[0]
    out = np.empty(n, dtype=np.int32)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)
    X, Eptr, O = xyz.ctypes.data_as(dp), e2n.ctypes.data_as(ip), out.ctypes.data_as(ip)
    assert lib.mgcfd_partition_rcb(n, X, 0, O) == -1                        # no parts
    assert lib.mgcfd_partition_rcb(n, None, 2, O) == -1
    assert lib.mgcfd_partition_rcb(n, X, 2, None) == -1
    assert lib.mgcfd_partition_rcb(0, None, 2, O) == 0                      # an empty set is a valid set
    assert lib.mgcfd_partition_graph(n, X, E, Eptr, tiny["base_array_index"], 2, b"kway", O) == 0
    assert set(out.tolist()) == {0, 1}
    assert lib.mgcfd_partition_graph(n, X, E, Eptr, tiny["base_array_index"] + 1, 2, b"kway", O) == -1      # an entry becomes -1
    assert lib.mgcfd_partition_graph(n, X, E, None, tiny["base_array_index"], 2, b"kway", O) == -1
    assert lib.mgcfd_partition_graph(n, X, E, Eptr, tiny["base_array_index"], 2, None, O) == -1
    assert lib.mgcfd_partition_graph(n, X, E, Eptr, tiny["base_array_index"], 2, b"metis5", O) == -1        # unknown method
    # more parts than nodes: every node still gets a valid owner
    few = np.ascontiguousarray(xyz[:3])
    o3 = np.empty(3, dtype=np.int32)
    assert lib.mgcfd_partition_rcb(3, few.ctypes.data_as(dp), 8, o3.ctypes.data_as(ip)) == 0
    assert ((o3 >= 0) & (o3 < 8)).all()
