"""One-launch ncu capture of the lane-change solve kernel in fixed-work mode (every instance runs exactly 30 Newton
steps: `MCPB200_DEFS=EXP_FIXED_STEPS=30`), so the capture carries its OWN Newton-step count: DRAM bytes and issued
warp-instructions per Newton step.  Writes profiles/r2_traffic.json (read by bench.py for `roofline.traffic`).

    python scripts/capture_traffic.py           (on a GPU box; needs ncu)
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, STEPS = 65536, 30
env = dict(os.environ, MCPB200_DEFS=f"EXP_FIXED_STEPS={STEPS}")
out = os.path.join(ROOT, "gpurun_out", "r2_traffic_ncu.csv")
os.makedirs(os.path.dirname(out), exist_ok=True)
cmd = ["ncu", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,"
       "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,lts__t_sector_hit_rate.pct",
       "--clock-control", "none", "-k", "regex:mcp_solve", "-s", "2", "-c", "1", "--csv", "--log-file", out,
       sys.executable, os.path.join(ROOT, "scripts", "prof_lane.py"), str(B)]
if "--parse-only" not in sys.argv:
    subprocess.run(cmd, env=env, check=True, cwd=ROOT, stdout=subprocess.DEVNULL)
vals = {}
rows = [r for r in csv.reader(open(out)) if len(r) > 5]
hdr = rows[0]
for r in rows[1:]:
    vals[r[hdr.index("Metric Name")]] = (float(r[hdr.index("Metric Value")].replace(",", "")), r[hdr.index("Metric Unit")])
unit_scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
dram = sum(vals[k][0] * unit_scale[vals[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
n = B * STEPS
rec = {"lane_change": {
    "dram_bytes_per_newton_step": dram / n,
    "warp_instructions_per_newton_step": vals["smsp__inst_executed.sum"][0] / n,
    "shared_wavefronts_per_newton_step": vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"][0] / n,
    "newton_steps_in_capture": n, "dram_bytes_in_capture": dram,
    "kernel_ms_under_ncu": vals["gpu__time_duration.sum"][0] * {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[vals["gpu__time_duration.sum"][1]],
    "l2_hit_rate_pct": vals["lts__t_sector_hit_rate.pct"][0],
    "source": f"profiles/r2_traffic.json: one ncu launch of mcp_solve_kernel, fixed-work mode (B = {B} instances × {STEPS} Newton "
              "steps each, MCPB200_DEFS=EXP_FIXED_STEPS=30, scripts/capture_traffic.py); (dram__bytes_read.sum + "
              "dram__bytes_write.sum) ÷ the launch's own Newton-step count"}}
with open(os.path.join(ROOT, "profiles", "r2_traffic.json"), "w") as f:
    json.dump(rec, f, indent=1)
print(json.dumps(rec, indent=1))
