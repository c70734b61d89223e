#!/usr/bin/env python
"""bench.py -- MG-CFD hot path on B200: flux-edge edges/s through full multigrid cycles.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mesh m6] [--variant owner]

One "step" is one multigrid V-cycle (euler3d.cpp:458-641) over the synthetic deck.  At N=1 the deck is
BASELINE.json configs[1] (Onera-M6-shaped, 4 levels).  Prints ONE JSON line (rank 0).

  value        flux-edge edge updates per second of whole-cycle time, deck resident in HBM, CUDA-event timed
  e2e          same metric through the C-ABI with HOST buffers: per step the flow state of every level is
               uploaded, one cycle runs, the state is fetched back (copies inside the timed region)
  roofline     compute_flux_edge_kernel alone: algorithmic bytes (32*E + 120*N per call) over
               the CUDA-event time of the flux launches inside the timed region, against the measured HBM peak
  cpu_baseline the reference's own elemental kernels (oracle/_ref, OpenMP block-coloured, all host threads)
               on a bounded sample of the same deck

`--impl reference` times that CPU implementation as the whole job instead (the reference has no GPU code in
its tree; its OP2 build cannot be produced offline).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner, for one) are sent to
# stderr by pointing fd 1 at fd 2; emit() writes the result line to the saved original stdout
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

RK = 3


def visits_per_cycle(n_levels):
    """level visits of one V-cycle: L0..L(n-1) upward then L(n-2)..L1 downward (euler3d.cpp:573-640)"""
    return list(range(n_levels)) + list(range(n_levels - 2, 0, -1)) if n_levels > 1 else [0]


def flux_edges_per_cycle(sizes):
    return RK * sum(sizes[l][1] for l in visits_per_cycle(len(sizes)))


def flux_bytes_per_cycle(sizes):
    """algorithmic bytes of the flux-edge loops of one cycle: 32*E + 120*N per invocation"""
    return RK * sum(32 * sizes[l][1] + 120 * sizes[l][0] for l in visits_per_cycle(len(sizes)))


def rk_stage_bytes_per_cycle(sizes):
    """algorithmic bytes of the fused Runge-Kutta stage launches of one cycle: per stage the flux-edge loop
    (32*E + 120*N) plus time_step (168*N); the last stage of a visit also does residual (120*N).  The boundary
    flux, calc_rms and count_bad_vals the kernel also performs are NOT counted (SURVEY.md 8d figures)."""
    return sum(RK * (32 * sizes[l][1] + 120 * sizes[l][0] + 168 * sizes[l][0]) + 120 * sizes[l][0]
               for l in visits_per_cycle(len(sizes)))


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i] == "Active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def reorder_for_cpu(levels0, perms, edge_orders):
    """Apply the GPU planner's locality ordering to a 0-based deck so that the CPU baseline is not
    handicapped by the shuffled file order."""
    out = []
    for l, lev in enumerate(levels0):
        new_of_old, eo = perms[l].astype(np.int64), edge_orders[l].astype(np.int64)
        old_of_new = np.empty_like(new_of_old)
        old_of_new[new_of_old] = np.arange(new_of_old.size)
        d = {
            "node_coordinates": lev["node_coordinates"][old_of_new],
            "edge-->node": new_of_old[lev["edge-->node"][eo]].astype(np.int32),
            "edge_weights": lev["edge_weights"][eo],
            "bnd_node-->node": new_of_old[lev["bnd_node-->node"]].astype(np.int32),
            "bnd_node-->group": lev["bnd_node-->group"],
            "bnd_node_weights": lev["bnd_node_weights"],
        }
        if "node-->mg_node" in lev:
            d["node-->mg_node"] = perms[l + 1].astype(np.int64)[lev["node-->mg_node"][old_of_new]].astype(np.int32)
        out.append(d)
    return out


def hilbert_orders(levels0):
    """node permutation + edge order per level without a GPU: same rule as the planner, via its numpy restatement."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import plan_oracle
    perms, orders = [], []
    for lev in levels0:
        p = plan_oracle.hilbert_renumber(lev["node_coordinates"])
        perms.append(p)
        orders.append(plan_oracle.sort_edges(lev["edge-->node"], p))
    return perms, orders


def cpu_run(levels0_ordered, n_cycles, warmup, threads):
    """oracle/_ref (the reference's own headers) if present, else the port; OpenMP block-coloured."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    flags = open("/proc/cpuinfo").read()
    fast_ok = " avx2" in flags and " fma" in flags
    for kind, label in (("ref_fast", "reference"), ("ref", "reference"), ("port_fast", "port"), ("port", "port")):
        if kind.endswith("_fast") and not fast_ok:
            continue
        if orc.available(kind):
            break
    o = orc.Oracle(kind)
    used = o.set_threads(threads)
    run = o.make_state(levels0_ordered)
    run.init()
    if warmup:
        run.run(warmup)
    rc, st = run.run(n_cycles)
    if rc != 0:
        raise RuntimeError(f"CPU baseline failed rc={rc}")
    return {"edges_per_s": st.flux_edges / st.wall_total, "cycles_per_s": n_cycles / st.wall_total,
            "flux_kernel_edges_per_s": st.flux_edges / st.wall_flux_edge, "wall_s": st.wall_total,
            "kind": label, "lib": kind, "cores": used}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh", default="m6")
    ap.add_argument("--variant", default="owner", choices=["owner", "emit", "gather", "colour", "atomic"])
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--cpu-cycles", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-fusion", action="store_true", help="one kernel per op_par_loop call site")
    ap.add_argument("--no-graphs", action="store_true", help="enqueue every launch instead of replaying CUDA graphs")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = the deck grows with N along x (default, the driver's scaling run); strong = the --mesh deck "
                         "itself is partitioned over the N GPUs (BASELINE configs[3])")
    ap.add_argument("--partitioner", default="geom", help="N>1: geom | kway | block | random (op_partition methods)")
    ap.add_argument("--transport", default="ipc", choices=["ipc", "nccl"],
                    help="N>1: direct peer stores over CUDA-IPC-mapped memory (default) or NCCL send/recv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    pkg = ge.load_package()
    slab = args.mesh in pkg.meshgen.SLAB_CONFIGS          # single-level decks that every rank generates for itself
    if slab and world > 1:
        # BASELINE configs[4]: the deck (150M nodes for rotor37_150m) is never materialised on one host; the job is a
        # fixed deck cut into x-slabs, i.e. strong scaling
        mesh, levels0 = None, None
        sizes = [pkg.meshgen.slab_sizes(args.mesh)]
        args.scaling = "strong"
    else:
        if slab:
            mesh = pkg.meshgen.make_slab_global(args.mesh)
        elif world > 1 and args.scaling == "weak":
            # weak scaling: the deck grows with the GPU count along x, so every rank's share stays about one --mesh deck
            mesh_name, dims, seed0 = pkg.meshgen.CONFIGS[args.mesh]
            dims = [(nx * world, ny, nz, None if ne is None else ne * world) for nx, ny, nz, ne in dims]
            mesh = pkg.meshgen.make_multigrid((mesh_name, dims, seed0))
        else:
            mesh = pkg.meshgen.make_multigrid(args.mesh)
        levels0 = [pkg.meshgen.zero_based(l) for l in mesh["levels"]]
        sizes = [(l["node_coordinates"].shape[0], l["edge-->node"].shape[0], l["bnd_node-->node"].shape[0]) for l in levels0]
    how = (" x%d along x (weak scaling" % world if args.scaling == "weak" else " partitioned over %d GPUs (strong scaling" % world) + \
        (", x-slabs generated per rank" if slab else f", {args.partitioner} partition") + ", 1 rank per GPU)"
    workload = (f"{args.mesh}{how if world > 1 else ''}"
                f": {len(sizes)}-level synthetic deck, nodes {[s[0] for s in sizes]}, edges {[s[1] for s in sizes]}; "
                f"step = 1 multigrid V-cycle (visits {visits_per_cycle(len(sizes))}, RK=3)")
    config = {"workload": workload, "mesh": args.mesh, "levels": len(sizes), "flux_variant": args.variant,
              "arith": "exact" if args.exact else "fast", "fused_schedule": args.variant in ("owner", "emit") and not args.no_fusion, "cuda_graphs": not args.no_graphs,
              "l2": "no flush between steps: the V-cycle working set (~%d MB) exceeds the 126 MB L2"
                    % (sum(300 * s[0] + 32 * s[1] for s in sizes) // 2**20)}
    nthreads = os.cpu_count() or 1

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        if levels0 is None:
            # a deck that is only ever generated per rank: the CPU arm runs rank 0's slab (owned + halo nodes, the edges
            # and boundary entries of its owned nodes) as a stand-alone deck -- a bounded sample of the workload
            d = pkg.meshgen.make_slab_rank(args.mesh, 0, world)
            levels0 = [{k: d[k] for k in ("node_coordinates", "edge-->node", "edge_weights", "bnd_node-->node", "bnd_node-->group",
                                          "bnd_node_weights")}]
            config["cpu_sample"] = f"x-slab of rank 0 of {world} ({d['node_coordinates'].shape[0]} nodes, {d['edge-->node'].shape[0]} edges)"
        perms, orders = hilbert_orders(levels0)
        ordered = reorder_for_cpu(levels0, perms, orders)
        r = cpu_run(ordered, args.steps, args.warmup, nthreads)
        sample = f"{args.steps} full V-cycles of the same deck (locality-renumbered), OpenMP block-coloured, {r['lib']}"
        line = {"impl": "reference", "metric": "mg_cycle_flux_edges_per_s", "value": r["edges_per_s"], "unit": "edges/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["wall_s"] / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "mg_cycles_per_s": r["cycles_per_s"],
                "cpu_baseline": {"value": r["edges_per_s"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"], "sample": sample},
                "e2e": {"value": r["edges_per_s"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if world > 1:
        if slab:
            lm = pkg.RankMesh(pkg.meshgen.make_slab_rank(args.mesh, rank, world))      # this rank's x-slab + halo planes only
        else:
            parts = pkg.partition_levels(mesh["levels"], mesh["base_array_index"], world, method=args.partitioner)
            lm = pkg.LocalMesh(mesh["levels"], mesh["base_array_index"], parts, rank, world)
        gpu = pkg.MGCFD(local_mesh=lm, device=local_rank, flux_variant=args.variant if args.variant == "emit" else "owner",
                        exact_arith=args.exact,
                        owner_chunk_nodes=args.chunk, graphs=not args.no_graphs)
        if args.transport == "ipc":
            mine = torch.frombuffer(bytearray(gpu.ipc_export()), dtype=torch.uint8).cuda()
            blobs = [torch.zeros(4096, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(blobs, mine)
            gpu.comm_init_ipc(b"".join(b.cpu().numpy().tobytes() for b in blobs))
            dist.barrier()
        else:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid.copy_(torch.frombuffer(bytearray(pkg.nccl_unique_id()), dtype=torch.uint8))
            dist.broadcast(uid, 0)
            gpu.comm_init_nccl(bytes(uid.cpu().numpy().tobytes()))
        config["transport"] = args.transport
        local_sizes = [(lm.sizes(l)[0], lm.sizes(l)[1], lm.sizes(l)[2]) for l in range(len(sizes))]
    else:
        gpu = pkg.MGCFD(mesh["levels"], base_array_index=mesh["base_array_index"], device=local_rank,
                        flux_variant=args.variant, exact_arith=args.exact, owner_chunk_nodes=args.chunk,
                        fuse=not args.no_fusion, graphs=not args.no_graphs)
        local_sizes = sizes
    stream = torch.cuda.ExternalStream(gpu.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # warm-up (also builds the flux plans and captures the one-cycle CUDA graphs)
    gpu.run_cycles(args.warmup + (args.warmup % 2))      # an even count leaves both one-cycle graphs captured
    launches0 = gpu.kernel_launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- timed region 1: K cycles exactly as a user runs them (graph replay), CUDA events on the library's stream
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    gpu.run_cycles(args.steps)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = gpu.kernel_launches() - launches0
    # ---- timed region 2: the same K cycles launch by launch, every flux-edge / fused-stage launch event-timed
    gpu.timers_enable(2)
    gpu.timers_reset()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    gpu.run_cycles(args.steps)
    e3.record(stream)
    barrier()
    ms_timed = max_over_ranks(e2.elapsed_time(e3))
    clocks = sampler.stop() if sampler else None
    fused = world > 1 or (args.variant in ("owner", "emit") and not args.no_fusion)
    flux_ms, flux_calls, flux_elems = gpu.timer("rk_stage" if fused else "compute_flux_edge")
    per_level = []
    for l in range(len(sizes)):
        ms_l, calls_l, _ = gpu.timer("rk_stage" if fused else "compute_flux_edge", l)
        per_level.append(round(1e3 * ms_l / max(calls_l, 1), 2))
    gpu.timers_enable(0)

    edges_step = flux_edges_per_cycle(sizes)          # edges of the whole (undecomposed) deck, cut edges counted once
    value = edges_step * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    # per-GPU roofline: this rank's launches move this rank's edges (owned + recomputed cut edges) and nodes
    flux_bytes = (rk_stage_bytes_per_cycle(local_sizes) if fused else flux_bytes_per_cycle(local_sizes)) * args.steps
    achieved = flux_bytes / (flux_ms * 1e-3) / 1e9
    kname = (f"flux_{args.variant if args.variant == 'emit' else 'owner'}_kernel<FUSE> = compute_flux_edge + compute_bnd_node_flux + time_step (+ residual) in one launch"
             if fused else f"compute_flux_edge_kernel[{args.variant}]")
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if fused and world == 1 and args.mesh == "m6" and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["dram_bytes_per_launch"]      # ncu --set full capture of the level-0 launch
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of the level-0 launch (profiles/r01_traffic.json)",
                "peak_source": peak_src,
                "algorithmic_bytes": ("per stage 32E+120N (flux-edge) + 168N (time_step), + 120N (residual) after the last stage"
                                      if fused else "32E+120N per launch"),
                "algorithmic_bytes_per_launch_L0": (32 * local_sizes[0][1] + 288 * local_sizes[0][0]) if fused else (32 * local_sizes[0][1] + 120 * local_sizes[0][0]),
                "avg_launch_us": 1e3 * flux_ms / max(flux_calls, 1), "launches": flux_calls,
                "kernel_edges_per_s": flux_elems / (flux_ms * 1e-3), "share_of_step": flux_ms / ms_timed,
                "ms_per_step_with_launch_timers": ms_timed / args.steps,
                "per_level_avg_launch_us": per_level,
                "note": "achieved = algorithmic bytes of all timed launches / their summed CUDA-event time; M6 levels are "
                        "L2-resident sized, see config.l2"}

    # end-to-end through the C-ABI with host buffers
    e2e = None
    if not args.no_e2e:
        pinned = [pkg.PinnedArray((s[0], 5)) for s in local_sizes]       # the caller's page-locked host buffers
        for l in range(len(sizes)):
            gpu.fetch_into(l, "variables", pinned[l].array)
        n_e2e = max(3, min(args.steps, 10))
        for it in range(2 + n_e2e):
            if it == 2:
                barrier()
                t0 = time.perf_counter()
            for l in range(len(sizes)):
                gpu.set(l, "variables", pinned[l].array)
            gpu.run_cycles(1)
            for l in range(len(sizes)):
                gpu.fetch_into(l, "variables", pinned[l].array)
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        for p_ in pinned:
            p_.free()
        nbytes = sum(s[0] * 40 for s in local_sizes) * world
        e2e = {"value": edges_step * n_e2e / dt, "unit": "edges/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": n_e2e, "ms_per_step": 1e3 * dt / n_e2e,
               "what": "per step: mgcfd_set_dat(variables) for every level from page-locked host arrays, mgcfd_run_cycles(1), "
                       "mgcfd_fetch_dat(variables) for every level back into them; wall clock around the loop"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        perms = [gpu.plan_query(l, "node_perm") for l in range(len(sizes))]
        orders = [gpu.plan_query(l, "edge_order") for l in range(len(sizes))]
        r = cpu_run(reorder_for_cpu(levels0, perms, orders), args.cpu_cycles, 1, nthreads)
        cpu = {"value": r["edges_per_s"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
               "sample": f"{args.cpu_cycles} full V-cycles (+1 warm-up) of the same deck with the GPU run's node/edge ordering, "
                         f"{r['lib']}, OpenMP block-coloured; flux kernel alone {r['flux_kernel_edges_per_s']:.3e} edges/s",
               "mg_cycles_per_s": r["cycles_per_s"]}
    halo_bytes = gpu.halo_bytes_sent()
    gpu.close()
    if rank == 0:
        config["halo_bytes_sent_rank0"] = halo_bytes
        line = {"metric": "mg_cycle_flux_edges_per_s", "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling if world > 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "mg_cycles_per_s": args.steps / (ms * 1e-3), "roofline": roofline, "cpu_baseline": cpu,
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
