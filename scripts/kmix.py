"""Compile the lane-change kernel (no GPU needed) and print registers + the SASS mix of the region between two markers."""
import subprocess, sys, os, glob
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcp_b200 import build as b
b.build_library()
from mcp_b200 import capi, problems
which = sys.argv[1] if len(sys.argv) > 1 else "lane"
m0, m1 = (sys.argv[2], sys.argv[3]) if len(sys.argv) > 3 else ("for (; left > 0; --left) {", "// ============ shared-memory window")
ir = {"lane": lambda: problems.lane_change_game().mcp.ir, "readme": lambda: problems.readme_qp().ir,
      "masked4": lambda: problems.masked_game(4, 30).mcp.ir, "qp": lambda: problems.random_qp(100, 100).ir}[which]()
h = capi.Handle(ir, capi.COMPILE_ONLY)
h.close()
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcp_b200", "_kcache")
f = max(glob.glob(os.path.join(root, "*.cu")), key=os.path.getmtime)
for fn in glob.glob(os.path.join(root, "*.cu")):
    if f"#define NX {ir.nx}\n" in open(fn).read(4000) and os.path.getmtime(fn) >= os.path.getmtime(f) - 1e9:
        pass
cub = f[:-3] + ".cubin"
out = subprocess.run(["cuobjdump", "-res-usage", cub], capture_output=True, text=True).stdout.splitlines()
for i, l in enumerate(out):
    if "Function mcp_" in l:
        print(l.strip(), out[i + 1].strip()[:60])
src = open(f).read().split("\n")
a = next(i for i, l in enumerate(src) if m0 in l) + 1
e = next(i for i, l in enumerate(src) if m1 in l and i > a) + 1
kern = sys.argv[4] if len(sys.argv) > 4 else "mcp_solve_kernel"
print(subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "sass_mix.py"), cub, kern, str(a), str(e - 3), f],
                     capture_output=True, text=True).stdout)
