"""Side bench of the `-b` single-cell mode at scale: a generated 10M-read BAM with CB:Z tags through the CLI, the
unmodified reference (oracle/_ref/regtools_ref) on a region of the same file beside it, outputs compared on that region.
    python tools/bench_barcodes.py [reads] [distinct_barcodes]        -> one JSON object on stdout"""
import hashlib, json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_bc = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
REGION = "chr1:1-40000000"
CLI = os.path.join(ROOT, "regtools_b200", "regtools")
REF = os.path.join(ROOT, "oracle", "_ref", "regtools_ref")
scratch = bench.SCRATCH
bam = os.path.join(scratch, f"rtjx_bc_{reads}_{n_bc}.bam")
if not os.path.exists(bam + ".bai"):
    subprocess.check_call([bench.BAMGEN, "gen", "--out", bam, "--config", "c2", "--reads", str(reads), "--seed", "1234", "--level", "6",
                           "--barcodes", str(n_bc)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def sha(p):
    return hashlib.sha256(open(p, "rb").read()).hexdigest()


def run(exe, tag, region=None):
    bed, bc = os.path.join(scratch, f"bc_{tag}.bed"), os.path.join(scratch, f"bc_{tag}.barcodes")
    cmd = [exe, "junctions", "extract", "-s", "XS", "-b", bc, "-o", bed] + (["-r", region] if region else []) + [bam]
    t0 = time.perf_counter()
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, bed, bc


out = {"workload": f"synthetic {reads}-read single-chrom BAM with CB:Z tags, {n_bc} distinct barcodes; regtools junctions extract -s XS -b",
       "bam_bytes": os.path.getsize(bam)}
run(CLI, "warm")                                     # page cache + driver warm-up
ts = [run(CLI, "ours")[0] for _ in range(3)]
_, bed, bc = run(CLI, "ours")
out["ours_whole_file_s"] = ts
out["ours_reads_per_s"] = reads / min(ts)
out["junction_lines"] = sum(1 for _ in open(bed))
out["barcode_file_bytes"] = os.path.getsize(bc)
t_r, bed_o, bc_o = run(CLI, "ours_region", REGION)
out["ours_region_s"] = t_r
if os.path.exists(REF):
    n_region, _ = bench.count_reads_reference(bam, REGION)
    t_ref, bed_r, bc_r = run(REF, "ref_region", REGION)
    out["reference_region"] = {"region": REGION, "reads": n_region, "seconds": t_ref, "reads_per_s": (n_region / t_ref) if n_region else None,
                               "bed12_identical": sha(bed_o) == sha(bed_r), "barcode_file_identical": sha(bc_o) == sha(bc_r)}
    if n_region:
        out["speedup_vs_reference_rate"] = out["ours_reads_per_s"] / (n_region / t_ref)
print(json.dumps(out))
