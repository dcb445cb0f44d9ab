"""Developer helper: times the device inflate of a whole synthetic BAM in one launch (rtjx_inflate_file).
    python tools/prof_inflate.py [reads] [check_md5: 0|1] [config]"""
import os, sys, time, zlib, gzip
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import regtools_b200 as rt
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cfg = sys.argv[3] if len(sys.argv) > 3 else "c2"
bam = bench.ensure_bam(cfg, reads, 6)
for i in range(2):
    ex = rt.JunctionsExtractor(bam, ".", 0)
    t0 = time.perf_counter(); data = ex.inflate_file(0); t1 = time.perf_counter()
    st = ex.stats(); ex.close()
    print(f"inflate_file {1e3*(t1-t0):.0f} ms wall; kernel {st['inflate_kernel_ms']:.2f} ms; {st['inflated_bytes']/1e9:.2f} GB out; {st['inflated_bytes']/st['inflate_kernel_ms']/1e6:.1f} GB/s")
if len(sys.argv) > 2 and sys.argv[2] == "1":
    import hashlib
    want = hashlib.md5(gzip.decompress(open(bam, 'rb').read())).hexdigest()
    print("md5 equal to zlib:", hashlib.md5(data).hexdigest() == want)
