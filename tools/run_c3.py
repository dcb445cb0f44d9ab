"""Developer helper: BASELINE configs[2] — synthetic 100M-read whole-genome BAM, 150 bp paired, 12% spliced.
Generates the BAM, runs ours end to end (device feeder) twice, checks device-vs-host-feeder equality of the
junction table, and times the unmodified reference on a bounded region of the same BAM."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import regtools_b200 as rt
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
t0 = time.perf_counter(); bam = bench.ensure_bam("c3", reads, 6); t_gen = time.perf_counter() - t0
size = os.path.getsize(bam)
res = {"reads": reads, "bam_bytes": size, "gen_s": round(t_gen, 1)}
tables = {}
for mode, name in ((0, "device"), (0, "device2"), (1, "host")):
    t0 = time.perf_counter()
    ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, inflate_mode=mode)
    ex.identify_junctions_from_BAM()
    tab = ex.junction_table()
    ex.output_file_ = f"/dev/shm/c3_{name}.bed"; ex.print_all_junctions()
    st = ex.stats(); ex.close()
    dt = time.perf_counter() - t0
    tables[name] = tab
    res[name] = {"wall_s": round(dt, 3), "reads_per_s": round(st["reads"] / dt), "reads": st["reads"], "junctions": len(tab),
                 "gpu_wait_s": round(st["host_wait_s"], 3), "staging_s": round(st["host_inflate_s"], 3), "grows": st["table_grows"]}
res["device_equals_host"] = bool(np.array_equal(tables["device"], tables["host"]))
region = "chr1:1-30000000"
n, ours_bed = bench.count_reads(bam, region, 0)
dt, kind, ref_bed = bench.time_reference(bam, region)
res["reference"] = {"kind": kind, "region": region, "reads": n, "wall_s": round(dt, 2), "reads_per_s": round(n / dt),
                    "bed12_identical_on_region": open(ours_bed).read() == open(ref_bed).read()}
res["speedup_e2e_vs_reference"] = round(res["device2"]["reads_per_s"] / res["reference"]["reads_per_s"], 1)
print(json.dumps(res))
