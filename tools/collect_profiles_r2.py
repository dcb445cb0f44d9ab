"""Developer helper: turns what tools/gpu_profiles_r2.sh brought back in gpurun_out/ into the tracked files under profiles/
(bench lines, launch-list table, reduced ncu metrics of every captured kernel, the roofline traffic figure, decoder inst/byte)."""
import collections, csv, json, os, re, shutil, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")


def cp(src, dst):
    s = os.path.join(G, src)
    if os.path.exists(s) and os.path.getsize(s) > 2:
        shutil.copyfile(s, os.path.join(P, dst)); print("copied", dst)


cp("r2_bench_c3.json", "r2_bench_c3_n1.json")
cp("r2_bench_c3_ref.json", "r2_bench_c3_n1_reference_arm.json")
cp("r2_regions_c4.json", "r2_regions_c4.json")
cp("r2_annotate.json", "r2_annotate.json")
cp("r2_barcodes.json", "r2_barcodes.json")
cp("r2_inflate_standalone.txt", "r2_inflate_standalone.txt")

# ---- launch list
L = os.path.join(G, "r2_launches.csv")
if os.path.exists(L):
    rows = [r for r in csv.reader(open(L, errors="replace")) if len(r) > 5]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    h = rows[hi]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1e-3)        # -> microseconds
        agg.setdefault(re.sub(r"\(.*", "", r[ki]), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(os.path.join(P, "r2_launches.md"), "w") as f:
        f.write("# ncu launch list (round 2) — `ncu --metrics gpu__time_duration.sum --clock-control none -c 1500` on `python bench.py --steps 2 --warmup 1 --no-courtesy` (C3: 100M-read BAM)\n\n")
        f.write("Per-launch times are cold-cache and SERIALISED by ncu (the device feeder's ten inflate streams do not overlap under the profiler): compare SHARES, not absolutes.  "
                "First 1500 launches of the run (warm-up step + the first timed step of the resident leg).\n\n| kernel | launches | avg µs | total ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| `{k}` | {len(v)} | {sum(v)/len(v):.1f} | {sum(v)/1000:.2f} | {100*sum(v)/tot:.2f}% |\n")
    shutil.copyfile(L, os.path.join(P, "r2_launches_raw.csv"))
    print("wrote r2_launches.md")

# ---- reduced ncu metrics
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_tag_requests.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
WHAT = {"r2_cigar_scan": "cigar_scan on a resident 30M-read C3 batch (tools/ab_scan.py 30000000 6 c3)",
        "r2_junction_merge": "junction_merge on the same batch",
        "r2_inflate_lanes": "bgzf_inflate_lanes on a 13M-read C3 file in ONE launch (tools/prof_inflate.py 13000000 0 c3: ~57k BGZF blocks, 3.72 GB out, 12 warps per SM)",
        "r2_feed_kernels": "device-feed kernels of one C2 pass (tools/prof_e2e.py 10000000 0 1)",
        "r2_annotate_kernel": "annotate_kernel, 2.03M junctions (tools/bench_annotate.py workload)"}
summ_path = os.path.join(P, "r2_ncu_full_summary.json")
summ = json.load(open(summ_path)) if os.path.exists(summ_path) else {}


def num(x):
    parts = x.split(" ")
    v = float(parts[0].replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}.get(parts[1] if len(parts) > 1 else "", 1)


for tag, what in WHAT.items():
    f = os.path.join(G, tag + ".raw.csv")
    if not os.path.exists(f) or os.path.getsize(f) < 1000:
        continue
    rr = list(csv.reader(open(f))); hh, units = rr[0], rr[1]
    ks = []
    for r in rr[2:]:
        d = {"kernel": r[hh.index("Kernel Name")].split("(")[0]}
        for w in WANT:
            if w in hh and r[hh.index(w)] not in ("", "n/a"):
                d[w] = r[hh.index(w)] + (" " + units[hh.index(w)] if units[hh.index(w)] else "")
        ks.append(d)
    summ[tag] = {"what": what, "kernels": ks}
    print("reduced", tag, len(ks), "kernels")
json.dump(summ, open(summ_path, "w"), indent=1)

if "r2_cigar_scan" in summ:
    sc = summ["r2_cigar_scan"]["kernels"][0]
    rd, wr = num(sc["dram__bytes_read.sum"]), num(sc["dram__bytes_write.sum"])
    json.dump({"cigar_scan_dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
               "source": "ncu --set full capture reduced to profiles/r2_ncu_full_summary.json (tools/gpu_profiles_r2.sh ncu)",
               "grid": sc.get("launch__grid_size"), "kernel": sc["kernel"], "algorithmic_read_bytes": 16.0 * 30000000 + 4.0 * 41847541,
               "workload": "30M-read C3 batch (150 bp paired, 12% spliced), resident; one launch"}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
if "r2_inflate_lanes" in summ:
    k = summ["r2_inflate_lanes"]["kernels"][0]
    out_bytes = None
    t = os.path.join(P, "r2_inflate_standalone.txt")
    if os.path.exists(t):
        m = re.findall(r"([0-9.]+) GB out", open(t).read())
        if m: out_bytes = float(m[-1]) * 1e9
    inst = num(k["smsp__inst_executed.sum"]); dur = num(k["gpu__time_duration.sum"])
    d = {"kernel": k["kernel"], "what": summ["r2_inflate_lanes"]["what"], "warp_instructions": inst, "duration_s": dur, "inflated_bytes": out_bytes,
         "warp_instructions_per_output_byte": inst / out_bytes if out_bytes else None, "inflated_GB_per_s_under_ncu": out_bytes / dur / 1e9 if out_bytes else None,
         "dram_bytes_read": num(k["dram__bytes_read.sum"]), "dram_bytes_written": num(k["dram__bytes_write.sum"]), "metrics": k}
    json.dump(d, open(os.path.join(P, "r2_inflate_lanes.json"), "w"), indent=1)
    print("wrote r2_inflate_lanes.json", d["warp_instructions_per_output_byte"], d["inflated_GB_per_s_under_ncu"])
