#!/usr/bin/env python
"""bench.py -- MG-CFD hot path on B200: flux-edge edges/s through full multigrid cycles.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mesh m6] [--variant owner]

One "step" is one multigrid V-cycle (euler3d.cpp:458-641) over the synthetic deck.  At EVERY N the deck is
BASELINE.json configs[3]: the Rotor37-shaped 8M-node 4-level deck, partitioned over the N GPUs (strong scaling;
it fits one GPU, so N=1 is the same job undecomposed).  The N=1 line also carries the Onera-M6-shaped cycle
(configs[1]) as the sub-record "m6".  Prints ONE JSON line (rank 0).

  value        flux-edge edge updates per second of whole-cycle time, deck resident in HBM, CUDA-event timed
  e2e          same metric through the C-ABI with HOST buffers: per step the flow state of every level is
               uploaded, one cycle runs, the state is fetched back (copies inside the timed region)
  roofline     compute_flux_edge_kernel alone: algorithmic bytes (32*E + 120*N per call) over
               the CUDA-event time of the flux launches inside the timed region, against the measured HBM peak
  cpu_baseline the reference's own elemental kernels (oracle/_ref, OpenMP block-coloured, all host threads)
               on a bounded sample of the same deck
  parity       after the timed regions the variables are re-initialised (euler3d.cpp:414-417), TWO cycles run on the
               GPU(s) and every level's owned-node state is compared with the CPU oracle's after the same cycles
               (<= 1e-10 relative to the variable's largest magnitude, north_star's tolerance; -v criterion of
               validation.h:46-100 counted beside it).  A mismatch makes the run fail (exit code 3).

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


def node_kernel_bytes_per_cycle(sizes):
    """algorithmic bytes per cycle of the fused node kernels (SURVEY.md 8d per-argument figures):
    visit_begin = copy_double (80 N) + calculate_dt (56 N) + get_min_dt (8 N) per visit; restrict = up_pre + up + up_post
    (48 N_fine + 216 N_coarse) per step up; down (148 N_fine + 64 N_coarse) per step down"""
    nl = len(sizes)
    n = [s[0] for s in sizes]
    return {"visit_begin": sum(144 * n[l] for l in visits_per_cycle(nl)),
            "restrict": sum(48 * n[l] + 216 * n[l + 1] for l in range(nl - 1)),
            "down": sum(148 * n[l] + 64 * n[l + 1] for l in range(nl - 1))}


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
                                          "-lms", "25", "-i", str(index)], stdout=subprocess.PIPE, text=True)
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


def cpu_info():
    """CPU model string, physical cores and hardware threads of this host."""
    model, phys = None, set()
    pid = cid = None
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name") and model is None:
                model = line.split(":", 1)[1].strip()
            elif line.startswith("physical id"):
                pid = line.split(":", 1)[1].strip()
            elif line.startswith("core id"):
                cid = line.split(":", 1)[1].strip()
            elif not line.strip():
                if pid is not None or cid is not None:
                    phys.add((pid, cid))
                pid = cid = None
    except OSError:
        pass
    return {"cpu_model": model, "physical_cores": len(phys) or None, "hardware_threads": os.cpu_count()}


def pick_oracle():
    """oracle/_ref (the reference's own headers compiled in place) if present, else the plain-C port"""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc
    flags = open("/proc/cpuinfo").read()
    fast_ok = " avx2" in flags and " fma" in flags
    for kind, label in (("ref_fast", "reference"), ("ref", "reference"), ("port_fast", "port"), ("port", "port")):
        if kind.endswith("_fast") and not fast_ok:
            continue
        if orc.available(kind):
            break
    return orc.Oracle(kind), kind, label


PARITY_CYCLES = 2      # cycles of the in-line parity check: both one-cycle CUDA graphs (the two parity states) get exercised


def cpu_run(levels0_ordered, n_cycles, warmup, threads, keep_first_cycle=False):
    """The CPU implementation of the path, OpenMP block-coloured on `threads` host threads: init, `warmup` cycles,
    then `n_cycles` timed cycles.  keep_first_cycle: also return every level's variables after the first PARITY_CYCLES
    cycles from the initial state (the parity reference; they count as warm-up)."""
    o, kind, label = pick_oracle()
    used = o.set_threads(threads)
    run = o.make_state(levels0_ordered)
    run.init()
    first = None
    if keep_first_cycle:
        rc, _ = run.run(PARITY_CYCLES)
        if rc != 0:
            raise RuntimeError(f"CPU oracle failed rc={rc}")
        first = [lv["var"].copy() for lv in run.levels]
        warmup = max(0, warmup - PARITY_CYCLES)
    if warmup:
        run.run(warmup)
    out = {"kind": label, "lib": kind, "cores": used, "first_cycle": first}
    if n_cycles > 0:
        rc, st = run.run(n_cycles)
        if rc != 0:
            raise RuntimeError(f"CPU baseline failed rc={rc}")
        out.update({"edges_per_s": st.flux_edges / st.wall_total, "cycles_per_s": n_cycles / st.wall_total,
                    "flux_kernel_edges_per_s": st.flux_edges / st.wall_flux_edge, "wall_s": st.wall_total})
    return out


def log(msg):
    sys.stderr.write(f"[bench {time.strftime('%H:%M:%S')}] {msg}\n")
    sys.stderr.flush()


def describe(mesh_name, sizes, world, scaling, slab, partitioner):
    how = ""
    if world > 1:
        how = (" x%d along x (weak scaling" % world if scaling == "weak" else " partitioned over %d GPUs (strong scaling" % world) + \
            (", x-slabs generated per rank" if slab else f", {partitioner} partition") + ", 1 rank per GPU)"
    return (f"{mesh_name}{how}: {len(sizes)}-level synthetic deck, nodes {[s[0] for s in sizes]}, edges {[s[1] for s in sizes]}; "
            f"step = 1 multigrid V-cycle (visits {visits_per_cycle(len(sizes))}, RK=3)")


def time_cycles(gpu, stream, steps, barrier, max_over_ranks):
    """K cycles exactly as a user runs them (CUDA-graph replay), CUDA events on the library's stream; ms for all K"""
    import torch
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    gpu.run_cycles(steps)
    e1.record(stream)
    barrier()
    return max_over_ranks(e0.elapsed_time(e1))


LOOPS = ("visit_begin", "compute_step_factor", "min_exchange", "rk_stage", "compute_flux_edge", "compute_bnd_node_flux", "time_step",
         "copy_double", "calculate_dt", "get_min_dt", "residual", "calc_rms", "count_bad_vals", "halo_wait", "halo_exchange",
         "restrict", "up_pre", "up", "up_post", "down")


def stage_timing(gpu, stream, steps, n_levels, fused, barrier, max_over_ranks, mode=2):
    """the same K cycles with the library's device timers on.  mode 2: every flux-edge / fused-stage launch bracketed by
    CUDA events, launch by launch (no graph).  mode 3: every call site timed INSIDE CUDA-graph replay (event-record nodes
    in the captured graph), which also gives the per-loop breakdown of a cycle as it actually runs."""
    import torch
    name = "rk_stage" if fused else "compute_flux_edge"
    gpu.timers_enable(mode)
    if mode == 3:
        gpu.run_cycles(2)          # capture the two timed graphs outside the measurement
    gpu.timers_reset()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(stream)
    gpu.run_cycles(steps)
    e3.record(stream)
    barrier()
    ms_timed = max_over_ranks(e2.elapsed_time(e3))
    flux_ms, flux_calls, flux_elems = gpu.timer(name)
    per_level, per_level_ms = [], []
    for l in range(n_levels):
        ms_l, calls_l, _ = gpu.timer(name, l)
        per_level.append(round(1e3 * ms_l / max(calls_l, 1), 2))
        per_level_ms.append(ms_l)
    loops = {}
    for ln in LOOPS:
        ms_l, calls_l, _ = gpu.timer(ln)
        if calls_l:
            loops[ln] = {"ms_per_step": ms_l / steps, "launch_sites_per_step": calls_l / steps}
    gpu.timers_enable(0)
    return {"ms_timed": ms_timed, "flux_ms": flux_ms, "calls": flux_calls, "elems": flux_elems, "per_level_us": per_level,
            "per_level_ms": per_level_ms, "loops": loops}


def m6_subrecord(pkg, device, steps, warmup, peak):
    """BASELINE configs[1]: the Onera-M6-shaped 4-level cycle on one GPU, reported inside the N=1 line"""
    import torch
    mesh = pkg.meshgen.make_multigrid("m6")
    sizes = [(l["node_coordinates"].shape[0], l["edge-->node"].shape[0], l["bnd_node-->node"].shape[0]) for l in mesh["levels"]]
    gpu = pkg.MGCFD(mesh["levels"], base_array_index=mesh["base_array_index"], device=device)
    stream = torch.cuda.ExternalStream(gpu.stream(), device=torch.device("cuda", device))
    sync = torch.cuda.synchronize
    gpu.run_cycles(warmup + warmup % 2)
    ms = time_cycles(gpu, stream, steps, sync, lambda x: x)
    st = stage_timing(gpu, stream, steps, len(sizes), True, sync, lambda x: x, mode=3)
    launches0 = gpu.kernel_launches()
    gpu.run_cycles(1)
    launches = gpu.kernel_launches() - launches0
    gpu.close()
    nbytes = rk_stage_bytes_per_cycle(sizes) * steps
    return {"workload": describe("m6", sizes, 1, "strong", False, "geom"), "steps": steps, "ms_per_step": ms / steps,
            "mg_cycles_per_s": steps / (ms * 1e-3), "edges_per_s": flux_edges_per_cycle(sizes) * steps / (ms * 1e-3),
            "launches_per_cycle": launches,
            "stage_frac_in_graph": nbytes / (st["flux_ms"] * 1e-3) / 1e9 / peak,
            "stage_per_level_us": st["per_level_us"],
            "cycle_breakdown_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in st["loops"].items()},
            "note": "levels are L2-sized (300K..81K nodes): the roofline fraction of the line is quoted on the 8M-node deck"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh", default="rotor37_8m", help="deck (default: BASELINE configs[3], the Rotor37-shaped 8M-node 4-level deck)")
    ap.add_argument("--variant", default="owner", choices=["owner", "emit", "gather", "colour", "atomic"])
    ap.add_argument("--exact", action="store_true")
    ap.add_argument("--chunk", type=int, default=64)
    ap.add_argument("--cpu-cycles", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-m6", action="store_true", help="N=1: skip the Onera-M6 sub-record (BASELINE configs[1])")
    ap.add_argument("--no-fusion", action="store_true", help="one kernel per op_par_loop call site")
    ap.add_argument("--no-graphs", action="store_true", help="enqueue every launch instead of replaying CUDA graphs")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N>1: strong = the --mesh deck itself is partitioned over the N GPUs (default, BASELINE configs[3]); "
                         "weak = the deck grows with N along x")
    ap.add_argument("--partitioner", default="geom", help="N>1: geom | kway | block | random (op_partition methods)")
    ap.add_argument("--transport", default="ipc", choices=["ipc", "nccl"],
                    help="N>1: direct peer stores over CUDA-IPC-mapped memory (default) or NCCL send/recv")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return                                        # the CPU arm is rank 0's job alone
    pkg = ge.load_package()
    t_start = time.time()
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
    if rank == 0:
        log(f"deck {args.mesh} ready after {time.time() - t_start:.1f} s")
    # `config` is the same object in both arms (the driver compares them); run-specific facts go into `run_info`
    config = {"workload": describe(args.mesh, sizes, world, args.scaling, slab, args.partitioner), "mesh": args.mesh,
              "levels": len(sizes), "scaling": args.scaling,
              "l2": "no flush between steps: one V-cycle streams ~%d MB, far beyond the 126 MB L2"
                    % (sum(300 * s[0] + 32 * s[1] for s in sizes) // 2**20)}
    nthreads = os.cpu_count() or 1
    cinfo = cpu_info()

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if levels0 is None:
            # a deck that is only ever generated per rank: the CPU arm runs rank 0's slab (owned + halo nodes, the edges
            # and boundary entries of its owned nodes) as a stand-alone deck -- a bounded sample of the workload
            d = pkg.meshgen.make_slab_rank(args.mesh, 0, world)
            levels0 = [{k: d[k] for k in ("node_coordinates", "edge-->node", "edge_weights", "bnd_node-->node", "bnd_node-->group",
                                          "bnd_node_weights")}]
            sample_what = f"x-slab of rank 0 of {world} ({d['node_coordinates'].shape[0]} nodes, {d['edge-->node'].shape[0]} edges)"
        else:
            sample_what = "the same deck"
        perms, orders = hilbert_orders(levels0)
        ordered = reorder_for_cpu(levels0, perms, orders)
        log(f"locality ordering done after {time.time() - t_start:.1f} s")
        r = cpu_run(ordered, args.steps, args.warmup, nthreads)
        sample = (f"{args.steps} full V-cycles of {sample_what} (locality-renumbered like the GPU run), OpenMP block-coloured, "
                  f"{r['lib']}, {r['cores']} threads")
        line = {"impl": "reference", "metric": "mg_cycle_flux_edges_per_s", "value": r["edges_per_s"], "unit": "edges/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * r["wall_s"] / args.steps,
                "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "mg_cycles_per_s": r["cycles_per_s"],
                "cpu_baseline": dict({"value": r["edges_per_s"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
                                      "sample": sample}, **cinfo),
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

    run_info = {"flux_variant": args.variant, "arith": "exact" if args.exact else "fast",
                "fused_schedule": world > 1 or (args.variant in ("owner", "emit") and not args.no_fusion),
                "cuda_graphs": not args.no_graphs}
    lm = None
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
        run_info["transport"] = args.transport
        local_sizes = [(lm.sizes(l)[0], lm.sizes(l)[1], lm.sizes(l)[2]) for l in range(len(sizes))]
    else:
        gpu = pkg.MGCFD(mesh["levels"], base_array_index=mesh["base_array_index"], device=local_rank,
                        flux_variant=args.variant, exact_arith=args.exact, owner_chunk_nodes=args.chunk,
                        fuse=not args.no_fusion, graphs=not args.no_graphs)
        local_sizes = sizes
    stream = torch.cuda.ExternalStream(gpu.stream(), device=torch.device("cuda", local_rank))
    if rank == 0:
        log(f"context planned after {time.time() - t_start:.1f} s")

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
    if rank == 0:
        log(f"warm-up done after {time.time() - t_start:.1f} s")
    launches0 = gpu.kernel_launches()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    # ---- timed region 1: K cycles exactly as a user runs them (graph replay), CUDA events on the library's stream
    ms = time_cycles(gpu, stream, args.steps, barrier, max_over_ranks)
    launches = gpu.kernel_launches() - launches0
    # ---- timed region 2: the same K cycles launch by launch, every flux-edge / fused-stage launch event-timed
    fused = run_info["fused_schedule"]
    st2 = stage_timing(gpu, stream, args.steps, len(sizes), fused, barrier, max_over_ranks, mode=2)
    # ---- timed region 3: the same K cycles as CUDA-graph replays with event-record nodes around every call site
    st = stage_timing(gpu, stream, args.steps, len(sizes), fused, barrier, max_over_ranks, mode=3) if not args.no_graphs else st2
    if not st["calls"]:
        st = st2
    clocks = sampler.stop() if sampler else None
    flux_ms, flux_calls, flux_elems, ms_timed = st["flux_ms"], st["calls"], st["elems"], st["ms_timed"]

    edges_step = flux_edges_per_cycle(sizes)          # edges of the whole (undecomposed) deck, cut edges counted once
    value = edges_step * args.steps / (ms * 1e-3)
    peak, peak_src = measured_peaks()
    # per-GPU roofline: this rank's launches move this rank's edges (owned + recomputed cut edges) and nodes
    flux_bytes = (rk_stage_bytes_per_cycle(local_sizes) if fused else flux_bytes_per_cycle(local_sizes)) * args.steps
    achieved = flux_bytes / (flux_ms * 1e-3) / 1e9
    kname = ("rk_stage2_kernel (exact build: flux_owner_kernel<FUSE>) = compute_flux_edge + compute_bnd_node_flux + time_step (+ residual) in one launch"
             if fused else f"compute_flux_edge_kernel[{args.variant}]")
    # level-0 launches alone, both accountings: the fused stage's own bytes (32E + 288N, +120N after the last stage) and the
    # flux-edge loop's (32E + 120N) over the SAME launch time -- the latter is north_star's ">= 60 % on the flux-edge loop"
    # read as strictly as possible (the launch also does time_step / residual, whose bytes are then not credited)
    E0, N0 = local_sizes[0][1], local_sizes[0][0]
    l0_calls = RK * args.steps
    l0_us = 1e3 * st["per_level_ms"][0] / max(l0_calls, 1)
    frac_fused_l0 = ((32 * E0 + 288 * N0 + 40 * N0) if fused else (32 * E0 + 120 * N0)) / (l0_us * 1e-6) / 1e9 / peak
    frac_flux_only_l0 = (32 * E0 + 120 * N0) / (l0_us * 1e-6) / 1e9 / peak
    traffic, traffic_note = None, "no ncu --set full capture of this deck committed yet"
    tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if fused and world == 1 and os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("mesh") == args.mesh:
            traffic, traffic_note = tj["dram_bytes_per_launch"], tj.get("note", "")
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": peak_src,
                "algorithmic_bytes": ("per stage 32E+120N (flux-edge) + 168N (time_step), + 120N (residual) after the last stage"
                                      if fused else "32E+120N per launch"),
                "algorithmic_bytes_per_launch_L0": (32 * E0 + 288 * N0) if fused else (32 * E0 + 120 * N0),
                "frac_fused_L0": frac_fused_l0, "frac_flux_only_L0": frac_flux_only_l0, "avg_launch_us_L0": l0_us,
                "avg_launch_us": 1e3 * flux_ms / max(flux_calls, 1), "launches": flux_calls,
                "kernel_edges_per_s": flux_elems / (flux_ms * 1e-3), "share_of_step": flux_ms / ms_timed,
                "timing": "CUDA events recorded by event-record nodes INSIDE the replayed one-cycle graphs (timers mode 3)" if st is not st2
                          else "CUDA events around every launch, launch by launch (timers mode 2)",
                "ms_per_step_with_timers": ms_timed / args.steps,
                "other_kernels_ms_per_step": (ms_timed - flux_ms) / args.steps,
                "per_level_avg_launch_us": st["per_level_us"],
                "cycle_breakdown_ms_per_step": {k: round(v["ms_per_step"], 5) for k, v in st["loops"].items()},
                "launch_timed": {"frac": flux_bytes / (st2["flux_ms"] * 1e-3) / 1e9 / peak, "per_level_avg_launch_us": st2["per_level_us"],
                                 "ms_per_step": st2["ms_timed"] / args.steps},
                "note": "achieved = algorithmic bytes of all timed stage launches / their summed CUDA-event time (every level of the "
                        "deck exceeds the 126 MB L2); frac_*_L0 = the level-0 launches alone under both accountings; "
                        "cycle_breakdown = device time per call site and cycle inside graph replay (its sum + gaps = ms_per_step_with_timers)"}
    # the other kernels of the cycle against the same roofline (timers mode 3 only: they are not timed launch by launch)
    nk = node_kernel_bytes_per_cycle(local_sizes)
    roofline["node_kernels"] = {k: {"ms_per_step": round(st["loops"][k]["ms_per_step"], 5), "algorithmic_mb_per_step": round(b / 1e6, 1),
                                    "frac": round(b / (st["loops"][k]["ms_per_step"] * 1e-3) / 1e9 / peak, 4)}
                                for k, b in nk.items() if k in st["loops"] and st["loops"][k]["ms_per_step"] > 0 and b > 0}
    roofline["node_kernels_note"] = ("algorithmic bytes of the reference's separate loops (visit_begin = copy_double + calculate_dt + get_min_dt; "
                                     "restrict = up_pre + up + up_post; down) over the fused kernel's in-graph time; a fraction above 1 means "
                                     "the fusion moves fewer bytes than the separate loops stream")
    if rank == 0:
        log(f"timed regions done after {time.time() - t_start:.1f} s: {ms / args.steps:.3f} ms/cycle, stage frac {achieved / peak:.3f}")

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
            arrs = [p_.array for p_ in pinned]
            gpu.run_cycles_host(1, arrs, arrs)       # upload every level, one cycle, fetch every level: one C-ABI call
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        for p_ in pinned:
            p_.free()
        nbytes = sum(s[0] * 40 for s in local_sizes)
        if world > 1:
            t = torch.tensor([float(nbytes)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            nbytes = int(t.item())
        e2e = {"value": edges_step * n_e2e / dt, "unit": "edges/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": nbytes, "steps": n_e2e, "ms_per_step": 1e3 * dt / n_e2e,
               "what": "per step one mgcfd_run_cycles_host(1, in, out): the flow state of every level goes up from page-locked host "
                       "arrays, one V-cycle runs, every level comes back into them (one GPU: coarse-level copies overlap the "
                       "level visits on a copy stream; N>1: set_dat / run_cycles / fetch_dat in sequence); wall clock around the loop"}
        if rank == 0:
            log(f"e2e done after {time.time() - t_start:.1f} s: {1e3 * dt / n_e2e:.2f} ms/step")

    # ---- CPU arm beside it (rank 0, N=1) and the parity reference (rank 0, every N): the oracle's state after ONE cycle
    cpu, ref_first = None, None
    want_parity = not args.no_parity and levels0 is not None
    if rank == 0 and (want_parity or (world == 1 and not args.no_cpu)):
        # the CPU runs the deck in the GPU planner's locality order (at N>1 from a planning-only context over the whole deck)
        pl = gpu if world == 1 else pkg.MGCFD(mesh["levels"], base_array_index=mesh["base_array_index"], device=-1, init=False)
        perms = [pl.plan_query(l, "node_perm") for l in range(len(sizes))]
        orders = [pl.plan_query(l, "edge_order") for l in range(len(sizes))]
        if world > 1:
            pl.close()
        deck = reorder_for_cpu(levels0, perms, orders)
        n_cpu = 0 if (world > 1 or args.no_cpu) else args.cpu_cycles
        r = cpu_run(deck, n_cpu, 1, nthreads, keep_first_cycle=True)
        ref_first = [r["first_cycle"][l][perms[l].astype(np.int64)] for l in range(len(sizes))]      # back to file order
        if n_cpu:
            cpu = dict({"value": r["edges_per_s"], "unit": "edges/s", "cores": r["cores"], "kind": r["kind"],
                        "sample": f"{n_cpu} full V-cycles (+{PARITY_CYCLES} warm-up) of the same deck with the GPU run's node/edge ordering, "
                                  f"{r['lib']}, OpenMP block-coloured; flux kernel alone {r['flux_kernel_edges_per_s']:.3e} edges/s",
                        "mg_cycles_per_s": r["cycles_per_s"]}, **cinfo)
        log(f"CPU oracle done after {time.time() - t_start:.1f} s")

    # ---- parity of the benchmarked configuration: re-initialise the variables, ONE cycle, compare with the oracle
    parity = {"checked": False, "why": "disabled (--no-parity)" if args.no_parity else "deck generated per rank: no single-host oracle run"}
    failed = False
    if want_parity:
        barrier()                        # rank 0 may have spent a while on the CPU oracle: start the cycle together
        gpu.reinit_variables()
        gpu.run_cycles(PARITY_CYCLES)
        max_rel, n_bad, vcount, bitwise = 0.0, 0, 0, True
        for l in range(len(sizes)):
            n_glob = sizes[l][0]
            if world > 1:
                # owned rows of this rank at their global (file) positions; the sum over ranks is the whole level
                got_local = gpu.fetch(l, "variables")
                gids = torch.from_numpy(lm.query(l, "global_node").astype(np.int64)[:gpu.n_owned[l]]).cuda()
                full = torch.zeros((n_glob, 5), dtype=torch.float64, device="cuda")
                full[gids] = torch.from_numpy(got_local[:gpu.n_owned[l]]).cuda()
                dist.all_reduce(full)
                got = full.cpu().numpy() if rank == 0 else None
                del full
                # -v criterion (validation.h:46-100) on this rank's owned nodes against the oracle's rows
                ref_t = torch.zeros((n_glob, 5), dtype=torch.float64, device="cuda")
                if rank == 0:
                    ref_t.copy_(torch.from_numpy(ref_first[l]))
                dist.broadcast(ref_t, 0)
                master = ref_t[torch.from_numpy(lm.query(l, "global_node").astype(np.int64)).cuda()].cpu().numpy()
                del ref_t
                c = torch.tensor([float(gpu.validate(l, master))], dtype=torch.float64, device="cuda")
                dist.all_reduce(c)
                vcount += int(c.item())
            else:
                got = gpu.fetch(l, "variables")
                vcount += gpu.validate(l, ref_first[l])
            if rank == 0:
                ref = ref_first[l]
                scale = np.abs(ref).max(axis=0).clip(1e-300)
                err = np.abs(got - ref).max(axis=0) / scale
                max_rel = max(max_rel, float(err.max()))
                n_bad += int((~np.isfinite(got)).sum())
                bitwise = bitwise and bool(np.array_equal(got, ref))
        if rank == 0:
            ok = max_rel <= 1e-10 and n_bad == 0 and vcount == 0
            parity = {"checked": True, "cycles": PARITY_CYCLES, "levels": len(sizes), "max_rel_err": max_rel, "tolerance": 1e-10,
                      "bit_identical": bitwise, "validate_count": vcount, "non_finite": n_bad, "ok": ok,
                      "oracle": r["lib"],
                      "what": f"variables re-initialised (euler3d.cpp:414-417), {PARITY_CYCLES} V-cycles on the benchmarked configuration, every "
                              f"level's owned-node state vs the CPU oracle after {PARITY_CYCLES} cycles; max over variables of |diff| / max|ref|; "
                              "validate_count = nodes failing the -v criterion of validation.h:46-100"}
            failed = not ok
            log(f"parity: max_rel_err {max_rel:.3e}, validate_count {vcount}, ok={ok}")

    halo_bytes = gpu.halo_bytes_sent()
    gpu.close()
    m6 = None
    if rank == 0 and world == 1 and not args.no_m6 and args.mesh != "m6":
        m6 = m6_subrecord(pkg, local_rank, 100, 4, peak)
        log(f"m6 sub-record done after {time.time() - t_start:.1f} s: {m6['ms_per_step']:.4f} ms/cycle")
    if rank == 0:
        run_info["halo_bytes_sent_rank0"] = halo_bytes
        line = {"metric": "mg_cycle_flux_edges_per_s", "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": config, "run_info": run_info, "mg_cycles_per_s": args.steps / (ms * 1e-3), "roofline": roofline,
                "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "m6": m6, "gpu_launches": launches, "clocks": clocks}
        emit(line)
    if world > 1:
        fl = torch.tensor([1.0 if failed else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(fl, op=dist.ReduceOp.MAX)
        failed = fl.item() > 0
        dist.destroy_process_group()
    if failed:
        sys.exit(3)


if __name__ == "__main__":
    main()
