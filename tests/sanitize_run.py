"""Small driver for compute-sanitizer (memcheck / racecheck / synccheck) on the B200:
    compute-sanitizer --tool racecheck python tests/sanitize_run.py
Runs every flux variant, both arithmetic builds, fused and call-site schedules, and a 3-rank group on a tiny deck."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402


def main():
    pkg = ge.load_package()
    mesh = pkg.meshgen.make_multigrid(("m6wing", [(9, 7, 5, 900), (5, 5, 3, 190)], 3))     # odd node counts: 315 / 75
    for variant in ("owner", "emit", "gather", "colour", "atomic"):
        for exact in (False, True):
            if variant == "emit" and exact:
                continue
            with pkg.MGCFD(mesh["levels"], flux_variant=variant, exact_arith=exact, graphs=False) as g:
                g.run_cycles(2)
                g.run_cycles_loopwise(1)
                g.unstructured_stream(0)
                g.sync()
    with pkg.MGCFD(mesh["levels"]) as g:          # graph replay (stage2 kernel with the folded step factor)
        g.run_cycles(3)
    with pkg.MGCFD(mesh["levels"], measure_mem_bound=True) as g:      # -b: stream kernel after every stage
        g.run_cycles(2)
    parts = pkg.partition_levels(mesh["levels"], 1, 3)
    lms = [pkg.LocalMesh(mesh["levels"], 1, parts, r, 3) for r in range(3)]
    ranks = [pkg.MGCFD(local_mesh=lm) for lm in lms]
    pkg.group_run_cycles(ranks, 2)
    for r in ranks:
        r.close()
    print("SANITIZE_RUN_DONE")


if __name__ == "__main__":
    main()
