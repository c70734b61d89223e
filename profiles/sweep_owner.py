#!/usr/bin/env python
"""Tuning sweep of the owner flux kernel (evidence tooling, not part of the product).

    python profiles/sweep_owner.py --mesh m6 --steps 20 \
        --configs "pipe=0;pipe=1;pipe=1,thr=192;pipe=1,chunk=128"

One process, the deck is generated once; each configuration creates its own context with the library's
experiment knobs set in the environment (MGCFD_OWNER_PIPE, MGCFD_OWNER_THREADS, MGCFD_OWNER_PIPE_CTAS,
MGCFD_OWNER_MAX_LOC, MGCFD_OWNER_MAX_EDGES), runs K multigrid cycles as a user does (CUDA-graph replay, CUDA
events on the library's stream) and then the same K cycles with every fused-stage launch event-timed.
Prints one line per configuration and appends a JSON record to gpurun_out/sweep.jsonl.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (helpers only; bench redirects fd 1 to stderr, emit() writes to the real stdout)
import __graft_entry__ as ge  # noqa: E402

KNOBS = {"pipe": "MGCFD_OWNER_PIPE", "thr": "MGCFD_OWNER_THREADS", "ctas": "MGCFD_OWNER_PIPE_CTAS", "minb": "MGCFD_OWNER_PIPE_MINB", "slot": "MGCFD_OWNER_SLOTTING", "lean": "MGCFD_OWNER_LEAN", "epi": "MGCFD_OWNER_EPILOGUE", "split": "MGCFD_OWNER_SLOT_SPLIT",
         "maxloc": "MGCFD_OWNER_MAX_LOC", "maxedges": "MGCFD_OWNER_MAX_EDGES",
         "pdl": "MGCFD_PDL", "s2": "MGCFD_STAGE2", "tiles": "MGCFD_STAGE2_TILES", "s2minb": "MGCFD_STAGE2_MINB", "pf": "MGCFD_STAGE2_PF"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="m6")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--configs", default="pipe=0;pipe=1")
    args = ap.parse_args()
    import torch
    pkg = ge.load_package()
    mesh = pkg.meshgen.make_multigrid(args.mesh)
    sizes = [(l["node_coordinates"].shape[0], l["edge-->node"].shape[0]) for l in mesh["levels"]]
    peak, _ = bench.measured_peaks()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for cfg in args.configs.split(";"):
        kv = dict(x.split("=") for x in cfg.split(",") if x)
        for k in KNOBS.values():
            os.environ.pop(k, None)
        for k, v in kv.items():
            if k in KNOBS:
                os.environ[KNOBS[k]] = v
        variant, chunk = kv.get("variant", "owner"), int(kv.get("chunk", 64))
        try:
            gpu = pkg.MGCFD(mesh["levels"], base_array_index=mesh["base_array_index"], flux_variant=variant,
                            owner_chunk_nodes=chunk)
            stream = torch.cuda.ExternalStream(gpu.stream())
            gpu.run_cycles(args.warmup + args.warmup % 2)
            torch.cuda.synchronize()
            sampler = bench.ClockSampler(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            gpu.run_cycles(args.steps)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            gpu.timers_enable(2)
            gpu.timers_reset()
            gpu.run_cycles(args.steps)
            torch.cuda.synchronize()
            flux_ms, calls, _ = gpu.timer("rk_stage")
            per_level = []
            for l in range(len(sizes)):
                ms_l, calls_l, _ = gpu.timer("rk_stage", l)
                per_level.append(round(1e3 * ms_l / max(calls_l, 1), 2))
            gpu.timers_enable(0)
            clocks = sampler.stop()
            power = [float(r[2]) for r in sampler.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
            clocks["power_w_max"] = max(power) if power else None
            sm_all = [float(r[0]) for r in sampler.rows if r and r[0].replace(".", "").isdigit()]
            clocks["sm_mhz_min"] = min(sm_all) if sm_all else None
            nbytes = bench.rk_stage_bytes_per_cycle([(s[0], s[1], 0) for s in sizes]) * args.steps
            frac = nbytes / (flux_ms * 1e-3) / 1e9 / peak
            stats = [int(x) for x in gpu.plan_query(0, "owner_stats")]
            checksum = float(abs(gpu.fetch(0, "variables")).sum())
            gpu.close()
            rec = {"mesh": args.mesh, "config": cfg, "ms_per_cycle": round(ms, 4), "frac": round(frac, 4),
                   "edges_per_s": bench.flux_edges_per_cycle([(s[0], s[1], 0) for s in sizes]) / (ms * 1e-3),
                   "per_level_launch_us": per_level, "owner_stats_L0": stats, "checksum_L0": checksum, "clocks": clocks}
        except Exception as ex:  # keep sweeping: a configuration that does not fit is a result too
            rec = {"mesh": args.mesh, "config": cfg, "error": repr(ex)}
        with open(os.path.join(ROOT, "gpurun_out", "sweep.jsonl"), "a") as f:
            f.write(json.dumps(rec) + "\n")
        bench.emit(rec)


if __name__ == "__main__":
    main()
