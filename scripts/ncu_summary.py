"""Summarise one `ncu --set full` report into profiles/<name>_ncu_full_summary.json and refresh profiles/traffic.json.

usage: python scripts/ncu_summary.py report.ncu-rep out.json "<kernel description>" "<command>" sweeps_per_launch [traffic_key]
"""
import csv
import json
import subprocess
import sys

rep, out_path, desc, cmd, sweeps = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
key = sys.argv[6] if len(sys.argv) > 6 else None
WANT = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__time_duration.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__pcsamp_sample_count",
        "smsp__pcsamp_warps_issue_stalled_barrier", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_selected", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
        "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_sleeping",
        "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_branch_resolving",
        "smsp__pcsamp_warps_issue_stalled_dispatch_stall", "smsp__pcsamp_warps_issue_stalled_no_instructions"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {}
for h, u, v in zip(hdr, units, vals):
    if h in WANT:
        m[h] = {"unit": u, "value": v}
SC = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tot = sum(float(m[k]["value"].replace(",", "")) * SC[m[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
res = {"kernel": desc, "command": cmd, "dram_bytes_per_sweep": tot / sweeps, "metrics": m}
json.dump(res, open(out_path, "w"), indent=1)
print(json.dumps({k: v["value"] + " " + v["unit"] for k, v in m.items()}, indent=1))
print("dram bytes per sweep", tot / sweeps)
if key:
    tr = json.load(open("profiles/traffic.json"))
    tr[key] = {"dram_bytes_per_sweep": tot / sweeps, "source": f"{out_path} ({sweeps} sweeps per launch)"}
    json.dump(tr, open("profiles/traffic.json", "w"), indent=1)
