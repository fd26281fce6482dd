"""Summarises one `ncu --set full` report holding every kernel of ONE MSM into profiles/<tag>_kernels_summary.txt
and refreshes profiles/traffic.json with the DRAM bytes of the accumulation PHASE (all its kernels: batched-affine
rounds + XYZZ accumulation).  usage: python tools/ncu_phase_summary.py <tag> <report.ncu-rep> <traffic-key>"""
import csv
import json
import os
import re
import subprocess
import sys

tag, rep, tkey = sys.argv[1], sys.argv[2], sys.argv[3]
# `rep` is either a .ncu-rep or the CSV of `ncu -i <rep> --page raw --csv` (the reports of 20 kernels exceed what
# gpurun copies back, so the CSV is made on the GPU box)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
hdr, units = rows[0], rows[1]
ui = dict(zip(hdr, units))
keep = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
phase = re.compile(r"k_ba_|k_accumulate|k_heavy")
tot_traffic, tot_ms, lines = 0.0, 0.0, []
for r in rows[2:]:
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("b200msm::", "")
    lines.append(f"\n== {name}")
    for k in keep:
        if k in d:
            lines.append(f"{k:86s} {d[k]:>18s} {ui[k]}")
    if phase.search(name):
        f = lambda k: float(d[k].replace(",", "")) * scale.get(ui[k], 1)
        tot_traffic += f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
        t = float(d["gpu__time_duration.sum"].replace(",", ""))
        tot_ms += {"ns": t / 1e6, "us": t / 1e3, "ms": t, "s": t * 1e3}.get(ui["gpu__time_duration.sum"], t / 1e6)
out = f"profiles/{tag}_kernels_summary.txt"
with open(out, "w") as f:
    f.write(f"# ncu --set full --clock-control none --import-source on; one MSM, every kernel of interest; source: {rep}\n")
    f.write(f"# accumulation phase (k_ba_* + k_accumulate* + k_heavy*): {tot_ms:.3f} ms under ncu, DRAM traffic {tot_traffic/1e9:.3f} GB\n")
    f.write("\n".join(lines) + "\n")
print(open(out).read()[:6000])
p = "profiles/traffic.json"
tj = json.load(open(p)) if os.path.exists(p) else {}
tj[tkey] = tot_traffic
tj["source"] = (f"profiles/{tag}_kernels_summary.txt: sum of dram__bytes_read+write over the kernels of the accumulation phase, one ncu --set full "
                "capture of ONE MSM (a recorded constant, refreshed with every kernel change; not measured in the bench run)")
json.dump(tj, open(p, "w"), indent=1)
