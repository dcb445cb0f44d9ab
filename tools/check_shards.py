"""Developer helper: contig shards of a generated C3 BAM on ONE GPU — every shard through the file-mode device feeder (fresh handle),
through the resident-file feeder and through the host feeder; tables compared field by field, union compared with the whole file.
    python tools/check_shards.py [reads] [world]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import regtools_b200 as rt
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 4
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 8
bam = bench.ensure_bam("c3", reads, 6)
F = ("tid", "start", "end", "thick_start", "thick_end", "read_count", "strand", "left_ok", "right_ok")


def run(r, mode, resident=False):
    ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, inflate_mode=mode, shard_rank=r, shard_world=world, n_threads=threads)
    ex._handle()
    if resident:
        ex.stage_bam()
    os.environ["RTJX_TRACE"] = "1"
    ex.identify_junctions_from_BAM()
    t = np.sort(ex.junction_table(), order=["tid", "start", "end", "strand"])
    st = ex.stats()
    ex.close()
    return t, st


whole = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000)
whole.identify_junctions_from_BAM()
n_whole = len(whole.junction_table()); whole.close()
tot = {"file": 0, "resident": 0}
for r in range(world):
    tf, sf = run(r, 2)
    tr, sr = run(r, 2, resident=True)
    same = len(tf) == len(tr) and all(np.array_equal(tf[f], tr[f]) for f in F)
    print(f"shard {r}/{world}: file-mode {len(tf)} junctions / {sf['reads']} reads / {sf['bgzf_blocks']} blocks; resident {len(tr)} / {sr['reads']} / {sr['bgzf_blocks']}; "
          f"equal={same}; device path taken: file {sf['host_parse_s'] == 0.0}, resident {sr['host_parse_s'] == 0.0}", flush=True)
    tot["file"] += len(tf); tot["resident"] += len(tr)
print(f"whole file: {n_whole} junctions; sum over shards: file-mode {tot['file']}, resident {tot['resident']}")
