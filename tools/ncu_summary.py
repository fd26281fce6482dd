"""Turns gpurun_out/{launches.csv, prof_*.ncu-rep} into the committed summaries under profiles/.
usage: python tools/ncu_summary.py <round-tag> [launches.csv] [report.ncu-rep] [traffic-key]"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

tag = sys.argv[1]
launches = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/launches.csv"
rep = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/prof_acc.ncu-rep"
tkey = sys.argv[4] if len(sys.argv) > 4 else "k_accumulate_g1_2^20"
os.makedirs("profiles", exist_ok=True)

if os.path.exists(launches):
    lines = [l for l in open(launches) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e6, "us": v / 1e3, "ms": v, "s": v * 1e3}.get(row["Metric Unit"], v / 1e6)
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")[:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(f"profiles/{tag}_launches.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
        f.write(f"# source: {launches}; total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")
        f.write(f"{'ms':>10s} {'launches':>8s} {'ms/launch':>10s} {'share':>6s}  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{t:10.3f} {n:8d} {t/n:10.4f} {100*t/tot:5.1f}%  {k}\n")
    print(open(f"profiles/{tag}_launches.txt").read())

if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = re.compile(r"^(Kernel Name|gpu__time_duration.sum|dram__bytes_(read|write).sum|launch__registers_per_thread|launch__occupancy_limit_registers|"
                      r"sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed|sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed|"
                      r"sm__inst_executed_pipe_(alu|fma).avg.pct_of_peak_sustained_active|sm__throughput.avg.pct_of_peak_sustained_elapsed|"
                      r"gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed|sm__warps_active.avg.pct_of_peak_sustained_active|"
                      r"smsp__issue_active.avg.per_cycle_active|smsp__inst_executed.sum|sm__cycles_elapsed.avg.per_second|"
                      r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|"
                      r"sm__inst_executed.sum|local_load|local_store|smsp__inst_executed_op_local.*sum)$")
    out = {}
    with open(f"profiles/{tag}_{os.path.basename(rep).replace('.ncu-rep','')}_summary.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source: {rep}\n")
        for r in rows[2:]:
            f.write("\n")
            for h, u, v in zip(hdr, units, r):
                if keep.match(h):
                    f.write(f"{h:88s} {v:>22s} {u}\n")
                    out[h] = v
        print(open(f.name).read())
    try:
        rd = float(out["dram__bytes_read.sum"].replace(",", ""))
        wr = float(out["dram__bytes_write.sum"].replace(",", ""))
        ui = dict(zip(hdr, units))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tr = rd * scale[ui["dram__bytes_read.sum"]] + wr * scale[ui["dram__bytes_write.sum"]]
        p = "profiles/traffic.json"
        d = json.load(open(p)) if os.path.exists(p) else {}
        d[tkey] = tr
        json.dump(d, open(p, "w"), indent=1)
        print("traffic", tkey, tr)
    except Exception as e:
        print("no traffic:", e)
