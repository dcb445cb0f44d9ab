"""Developer helper: end-to-end passes (rtjx_run + BED12) with wall-clock breakdown.
    python tools/prof_e2e.py [reads] [inflate_mode] [reps] [config] [resident]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import regtools_b200 as rt
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
cfg = sys.argv[4] if len(sys.argv) > 4 else "c2"
resident = len(sys.argv) > 5 and sys.argv[5] == "resident"
bam = bench.ensure_bam(cfg, reads, 6)
keep = None
for i in range(reps):
    t0 = time.perf_counter()
    if resident and keep is not None:
        ex = keep; ex.clear(); ex.reset_stats()
    else:
        ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, inflate_mode=mode, profile=True)
        ex._handle()
        if resident:
            ex.stage_bam(); keep = ex
    t1 = time.perf_counter()
    ex.identify_junctions_from_BAM(); t2 = time.perf_counter()
    n = len(ex.junction_table()); t3 = time.perf_counter()
    ex.output_file_ = "/dev/shm/prof_e2e.bed"; ex.print_all_junctions(); t4 = time.perf_counter()
    st = ex.stats()
    if not resident:
        ex.close()
    t5 = time.perf_counter()
    print(f"rep {i}: create {1e3*(t1-t0):.1f} run {1e3*(t2-t1):.1f} finalize {1e3*(t3-t2):.1f} print {1e3*(t4-t3):.1f} close {1e3*(t5-t4):.1f} total {1e3*(t5-t0):.1f} ms; junctions {n}; {st['reads']/(t4-t1)/1e6:.1f} M reads/s")
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
import hashlib
print("bed12 sha256", hashlib.sha256(open("/dev/shm/prof_e2e.bed", "rb").read()).hexdigest())
