#!/usr/bin/env python
"""Run a few un-graphed multigrid cycles on one deck so that ncu can capture the fused-stage launches
(evidence tooling).  Usage under ncu:  ncu --set full -k regex:'rk_stage2|flux_owner' -s N -c M python profiles/ncu_stage.py --mesh rotor37_1m
Knobs: the library's environment variables (MGCFD_STAGE2, MGCFD_OWNER_LEAN, ...)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default="rotor37_1m")
    ap.add_argument("--cycles", type=int, default=2)
    ap.add_argument("--levels", type=int, default=0, help="keep only the first n levels of the deck (0 = all)")
    args = ap.parse_args()
    pkg = ge.load_package()
    mesh = pkg.meshgen.make_multigrid(args.mesh)
    levels = mesh["levels"][:args.levels] if args.levels else mesh["levels"]
    if args.levels:
        levels = [dict(l) for l in levels]
        levels[-1].pop("node-->mg_node", None)
    with pkg.MGCFD(levels, base_array_index=mesh["base_array_index"], graphs=False) as g:
        g.run_cycles(args.cycles)
        g.sync()


if __name__ == "__main__":
    main()
