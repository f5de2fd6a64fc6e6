#!/usr/bin/env python
"""Summarise an ncu capture of the dominant kernel: python profiles/scripts/ncu_summary.py <file.ncu-rep> <out.txt>
[--traffic-json profiles/push_traffic.json --workload c3_n10M_nnz100M --kernel-label "..."].
Reads the report with `ncu -i <rep> --page raw --csv` (no GPU needed) and writes the metrics the roofline argument uses:
duration, DRAM bytes, L2 (LTS) sector operations by source, L1 data-pipe wavefronts, issue utilisation, stall reasons.
With --traffic-json the DRAM bytes per launch are stored where bench.py looks them up (roofline.traffic)."""
import argparse
import csv
import json
import subprocess
import sys

ap = argparse.ArgumentParser()
ap.add_argument("rep")
ap.add_argument("out")
ap.add_argument("--traffic-json")
ap.add_argument("--workload", default="c3_n10M_nnz100M")
ap.add_argument("--kernel-label", default=None)
args = ap.parse_args()

raw = subprocess.run(["ncu", "-i", args.rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
WANT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "lts__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors.sum", "lts__t_sectors_srcunit_tex.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
        "lts__t_sectors_lookup_miss.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers"]
lines = [f"# ncu summary of {args.rep} (ncu -i ... --page raw --csv; captured with --set full --clock-control none "
         f"--cache-control none after warm-up launches)"]
last = None
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    last = d
    lines.append(f"kernel: {d.get('Kernel Name', '')}")
    for w in WANT:
        if w in d:
            lines.append(f"  {w:78s} {d[w]:>20s} {units[hdr.index(w)]}")
    try:
        sectors = float(d["lts__t_sectors.sum"].replace(",", ""))
        ltsc = float(d["lts__cycles_elapsed.avg"].replace(",", ""))
        lines.append(f"  derived: L2 sector operations per L2 clock under ncu = {sectors / ltsc:.1f} (chip cap ~197 = 6300 B/clk)")
    except Exception:
        pass
    st = [(h, float(d[h].replace(",", ""))) for h in hdr
          if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d[h]]
    for h, v in sorted(st, key=lambda t: -t[1])[:8]:
        lines.append(f"  {h:90s} {v:8.2f}")
open(args.out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
if args.traffic_json and last:
    try:
        tj = json.load(open(args.traffic_json))
    except Exception:
        tj = {}
    def num(k):
        v, u = float(last[k].replace(",", "")), units[hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    prev = tj.get(args.workload)
    entry = {"kernel": args.kernel_label or last.get("Kernel Name", ""),
             "dram_bytes_per_launch": int(num("dram__bytes_read.sum") + num("dram__bytes_write.sum")),
             "lts_sector_ops_per_launch": int(float(last["lts__t_sectors.sum"].replace(",", ""))),
             "source": f"{args.out} (written by profiles/scripts/ncu_summary.py from {args.rep}: ncu --set full, "
                       f"dram__bytes_read.sum + dram__bytes_write.sum of one warm launch)"}
    if prev:
        entry["previous"] = [{k: v for k, v in prev.items() if k != "previous"}] + prev.get("previous", [])
    tj[args.workload] = entry
    json.dump(tj, open(args.traffic_json, "w"), indent=1)
