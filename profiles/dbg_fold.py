import sys
sys.path.insert(0,'/root/repo')
import __graft_entry__ as ge
pkg = ge.load_package()
mesh = pkg.meshgen.make_multigrid("m6")
with pkg.MGCFD(mesh["levels"], base_array_index=mesh["base_array_index"]) as g:
    g.run_cycles(2)
    l0 = g.kernel_launches(); g.run_cycles(2); print("launches/cycle", (g.kernel_launches()-l0)/2)
