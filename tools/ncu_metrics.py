"""Developer helper: prints selected raw metrics of every kernel in an .ncu-rep.   python tools/ncu_metrics.py rep [regex]"""
import csv, re, subprocess, sys
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
# `rep` is an .ncu-rep, or the CSV of its raw page made on the GPU box (`ncu -i x.ncu-rep --page raw --csv > x.raw.csv`)
raw = open(rep).read() if rep.endswith(".csv") else subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); hh = rr[0]; units = rr[1]
DEFAULT = r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|dram__throughput.avg.pct|sm__warps_active.avg.pct|smsp__issue_active.avg.pct|smsp__inst_executed.sum$|launch__registers|launch__occupancy_limit|launch__grid_size|launch__block_size|smsp__average_warps?_issue_stalled.*_per_issue_active|smsp__average_warp_latency_issue_stalled|sm__inst_executed_pipe_lsu|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__thread_inst_executed_per_inst_executed|sm__throughput.avg.pct|lts__t_sector_hit_rate.pct|l1tex__t_sector_hit_rate.pct|smsp__warps_eligible.avg.per_cycle_active|smsp__warp_issue_stalled_.*_per_warp_active.pct"
pat = pat or re.compile(DEFAULT)
for r in rr[2:]:
    print("==", r[hh.index("Kernel Name")][:90])
    for i, h in enumerate(hh):
        if pat.search(h) and r[i] not in ("", "0", "n/a"):
            print(f"   {h} = {r[i]} {units[i]}")
