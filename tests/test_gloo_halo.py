"""World-size-2 run of the decomposed multigrid schedule on CPU over gloo.

The product's partitioner (host code of libmgcfd_b200.so) builds each rank's local mesh and halo lists; the
arithmetic is done by the CPU oracle on the local mesh; halos travel through torch.distributed (gloo) exactly where
the GPU schedule exchanges them (csrc/cycle.cu: variables after every Runge-Kutta stage / restrict / prolong,
residuals after a visit, min_dt all-reduce).  The assembled result must equal the undecomposed oracle run BIT FOR
BIT -- which pins the owner/halo/export/import index sets and the exchange points without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exchange(lm, l, arr, n_owned, rank):
    nbr, ep, ei, ip = (lm.query(l, k) for k in ("neighbour_rank", "export_ptr", "export_idx", "import_ptr"))
    reqs, recvs = [], []
    for k, q in enumerate(nbr):
        send = torch.from_numpy(np.ascontiguousarray(arr[ei[ep[k]:ep[k + 1]]]))
        recv = torch.empty((int(ip[k + 1] - ip[k]), arr.shape[1]), dtype=torch.float64)
        if send.numel():
            reqs.append(dist.isend(send, int(q)))
        if recv.numel():
            reqs.append(dist.irecv(recv, int(q)))
        recvs.append((k, recv))
    for r in reqs:
        r.wait()
    for k, recv in recvs:
        arr[n_owned + ip[k]:n_owned + ip[k + 1]] = recv.numpy()


def _worker(rank, world, port, name, cycles, out_dir, method="geom"):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import __graft_entry__ as ge
    import orc
    pkg = ge.load_package()
    mesh = pkg.meshgen.make_multigrid(name)
    parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], world, method=method)
    lm = pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, rank, world)
    o = orc.Oracle("port")
    nl = len(mesh["levels"])
    lev = []
    for l in range(nl):
        v = lm.level(l).contents
        n, E, B = v.n_nodes, v.n_edges, v.n_bnd_nodes
        as_np = lambda p, shape, dt: np.ctypeslib.as_array(p, shape=shape).astype(dt).copy()
        d = {"node_coordinates": as_np(v.node_coordinates, (n, 3), np.float64),
             "edge-->node": as_np(v.edge_to_node, (E, 2), np.int32), "edge_weights": as_np(v.edge_weights, (E, 3), np.float64),
             "bnd_node-->node": as_np(v.bnd_node_to_node, (B,), np.int32) if B else np.zeros(0, np.int32),
             "bnd_node-->group": as_np(v.bnd_node_to_group, (B,), np.int32) if B else np.zeros(0, np.int32),
             "bnd_node_weights": as_np(v.bnd_node_weights, (B, 3), np.float64) if B else np.zeros((0, 3))}
        if l + 1 < nl:
            d["node-->mg_node"] = np.maximum(as_np(v.node_to_mg_node, (n,), np.int32), 0)   # -1 entries are filtered below
            d["_mg_raw"] = as_np(v.node_to_mg_node, (n,), np.int32)
        lev.append(d)
    run = o.make_state(lev)
    run.init()
    A = run.levels
    no = [lm.sizes(l)[3] for l in range(nl)]
    for l in range(nl):
        _exchange(lm, l, A[l]["var"], no[l], rank)
    level, mg_dir, i = 0, 0, 0
    while i < cycles:
        a, n = A[level], no[level]
        o.copy_double(a["var"][:n], a["old"][:n])
        o.calculate_dt(a["var"][:n], a["vol"][:n], a["sf"][:n])
        m = torch.tensor([o.get_min_dt(a["sf"][:n])], dtype=torch.float64)
        dist.all_reduce(m, op=dist.ReduceOp.MIN)
        o.compute_step_factor(a["var"][:n], a["vol"][:n], float(m.item()), a["sf"][:n])
        for rk in range(3):
            o.compute_flux_edge(a["e2n"], a["var"], a["ewt"], a["flux"])
            o.compute_bnd_node_flux(a["bgroup"], a["bwt"], a["b2n"], a["var"], a["flux"])
            o.time_step(rk, a["sf"][:n], a["flux"][:n], a["old"][:n], a["var"][:n])
            a["flux"][n:] = 0.0                                   # increments that landed on halo nodes are discarded
            _exchange(lm, level, a["var"], n, rank)
        o.residual(a["old"][:n], a["var"][:n], a["res"][:n])
        if level >= 1:
            _exchange(lm, level, a["res"], n, rank)
        if nl <= 1:
            i += 1
        elif mg_dir == 0:
            level += 1
            fine, coarse = A[level - 1], A[level]
            raw = lev[level - 1]["_mg_raw"]
            gn = lm.query(level - 1, "global_node")
            keep = np.nonzero((raw >= 0) & (raw < no[level]))[0]   # children (owned or halo) of OWNED coarse nodes
            keep = keep[np.argsort(gn[keep], kind="stable")]       # in file order of the undecomposed mesh
            mgk = np.ascontiguousarray(raw[keep])
            vk = np.ascontiguousarray(fine["var"][keep])
            o.up_pre(mgk, coarse["var"], coarse["up_scratch"])
            o.up(mgk, vk, coarse["var"], coarse["up_scratch"])
            o.up_post(coarse["var"][:no[level]], coarse["up_scratch"][:no[level]])
            _exchange(lm, level, coarse["var"], no[level], rank)
            if level == nl - 1:
                mg_dir = 1
        else:
            level -= 1
            f, c, n = A[level], A[level + 1], no[level]
            o.down(np.ascontiguousarray(lev[level]["_mg_raw"][:n]), f["var"][:n], f["res"][:n], f["coords"][:n], c["res"], c["coords"])
            _exchange(lm, level, f["var"], n, rank)
            if level == 0:
                mg_dir, i = 0, i + 1
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"),
             **{f"var_L{l}": A[l]["var"][:no[l]] for l in range(nl)},
             **{f"gn_L{l}": lm.query(l, "global_node")[:no[l]] for l in range(nl)})
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name,cycles,world,method", [("tiny", 3, 2, "geom"), ("small", 2, 2, "geom"), ("small", 2, 3, "kway"),
                                                      ("tiny", 2, 5, "random")])
def test_decomposed_run_equals_undecomposed(tmp_path, name, cycles, world, method, meshgen, oracle_port):
    """2 ranks with the geometric partition, 3 ranks with the k-way graph partition, 5 ranks with the random one (every
    rank neighbours every other, halo lists at their worst): all bit-identical to the undecomposed run"""
    port = 29600 + (os.getpid() % 300) + 7 * world
    mp.spawn(_worker, args=(world, port, name, cycles, str(tmp_path), method), nprocs=world, join=True)
    lev0 = [meshgen.zero_based(l) for l in meshgen.make_multigrid(name)["levels"]]
    ref = oracle_port.make_state(lev0)
    ref.init()
    assert ref.run(cycles)[0] == 0
    for l in range(len(lev0)):
        full = np.full_like(ref.levels[l]["var"], np.nan)
        for r in range(world):
            z = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
            full[z[f"gn_L{l}"]] = z[f"var_L{l}"]
        assert np.array_equal(full, ref.levels[l]["var"]), (l, np.nanmax(np.abs(full - ref.levels[l]["var"])))
