"""Developer helper: turns gpurun_out/*.csv / *.ncu-rep into the tracked summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]                       # e.g. r1
launches = sys.argv[2]                  # launch-list csv
rep = sys.argv[3]                       # .ncu-rep with --set full
out = os.path.join(root, "profiles")

rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(out, f"{tag}_launches.md"), "w") as f:
    f.write(f"# ncu launch list ({tag}) — `ncu --metrics gpu__time_duration.sum --clock-control none` on `python bench.py --steps 2 --warmup 1`\n\n")
    f.write("Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.\n\n| kernel | launches | avg µs | total µs | share |\n|---|---:|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1000:.1f} | {sum(v)/1000:.1f} | {100*sum(v)/tot:.1f}% |\n")

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); hh = rr[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
units = rr[1]
summ = []
for r in rr[2:]:
    d = {w: r[hh.index(w)] for w in want if w in hh}
    d["units"] = {w: units[hh.index(w)] for w in want if w in hh and units[hh.index(w)]}
    summ.append(d)
json.dump(summ, open(os.path.join(out, f"{tag}_ncu_full_summary.json"), "w"), indent=1)
scan = [d for d in summ if "cigar_scan" in d["Kernel Name"]]
scan.sort(key=lambda d: -int(d.get("launch__grid_size", "0").replace(",", "")))          # the 10M-read resident batch, not an e2e slice
for d in scan[:1]:
    if "cigar_scan" in d["Kernel Name"]:
        rd = float(d["dram__bytes_read.sum"]) * (1e6 if d["units"]["dram__bytes_read.sum"] == "Mbyte" else 1e3 if d["units"]["dram__bytes_read.sum"] == "Kbyte" else 1)
        wr = float(d["dram__bytes_write.sum"]) * (1e6 if d["units"]["dram__bytes_write.sum"] == "Mbyte" else 1e3 if d["units"]["dram__bytes_write.sum"] == "Kbyte" else 1)
        json.dump({"cigar_scan_dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "source": os.path.basename(rep),
                   "grid": d.get("launch__grid_size"), "kernel": d["Kernel Name"].split("(")[0],
                   "workload": "10M-read C2 batch, resident"}, open(os.path.join(out, "roofline_traffic.json"), "w"), indent=1)
        break
print(open(os.path.join(out, f"{tag}_launches.md")).read())
for d in summ: print({k: v for k, v in d.items() if k != "units"})
