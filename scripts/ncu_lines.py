"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump by CUDA source line."""
import csv, sys, collections
def num(v):
    try: return int(float(v))
    except Exception: return 0
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if len(r) > 5 and r[0] == "Line No")
hdr = rows[hi]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = collections.OrderedDict()
cur_line, cur_src = None, ""
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    if r[0] == 'Line No': break   # a second kernel in the same report
    if r[0]: cur_line, cur_src = int(r[0]), r[1]
    if not r[2]: continue
    a = agg.setdefault(cur_line, [0, 0, cur_src, collections.Counter()])
    a[0] += num(r[si]); a[1] += num(r[ii])
    for i, h in stall_cols:
        if r[i]: a[3][h] += num(r[i])
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print("total samples", tot, "instructions", toti)
allst = collections.Counter()
for a in agg.values(): allst.update(a[3])
print("stall mix:", ", ".join(f"{k[6:]} {100*v/tot:.1f}%" for k, v in allst.most_common(8)))
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = ",".join(f"{k[6:]}:{100*v/max(a[0],1):.0f}" for k, v in a[3].most_common(3))
    print(f"{100*a[0]/tot:5.1f}% smp {100*a[1]/toti:5.1f}% inst L{line}: {a[2].strip()[:90]}  [{st}]")
# region summary: python ncu_lines.py dump.csv N source.cu
if len(sys.argv) > 3:
    src = open(sys.argv[3]).read().split("\n")
    marks = []
    for i, l in enumerate(src, 1):
        for key in ("void mcp_eval_newton_p0(", "void mcp_eval_sens_p0(", "double opval(", "void assemble_matrix(", "int band_solve(",
                    "// ============ register-resident window", "// ============ shared-memory window", "// ---- back substitution",
                    "double ftb_linesearch(", "mcp_solve_kernel(const SolveParams", "mcp_sens_kernel(const SensParams",
                    "// ---- pivot search", "// ---- publish the pivot row", "// ---- retire", "// ---- eliminate column",
                    "// ---- the entering row", "// ---- control: advance my instance", "// ---- one Newton step"):
            if key in l: marks.append((i, key))
    marks.sort()
    reg = collections.Counter(); regi = collections.Counter()
    for line, a in agg.items():
        name = "prologue"
        for ml, key in marks:
            if line >= ml: name = key
        reg[name] += a[0]; regi[name] += a[1]
    print("--- regions ---")
    for k, v in reg.most_common(): print(f"{100*v/tot:5.1f}% smp {100*regi[k]/toti:5.1f}% inst  {k}")
