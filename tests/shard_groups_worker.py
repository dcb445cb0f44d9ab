"""Worker of tests/test_gpu_device_feed.py::test_contig_shards_with_many_small_groups (its own process: the group size is read
from the environment once per process).  Every contig shard of a generated BAM, for several world sizes: the device feeder in
file mode, the device feeder on the file resident in HBM, and the host feeder must see the same alignments and build the same
table; the shards of a world must add up to the whole file."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import regtools_b200 as rt

bam = sys.argv[1]
worlds = [int(x) for x in sys.argv[2].split(",")]
F = ("tid", "start", "end", "thick_start", "thick_end", "read_count", "strand", "left_ok", "right_ok")


def run(mode, r, w, resident=False):
    ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, inflate_mode=mode, shard_rank=r, shard_world=w, n_threads=4)
    if resident:
        ex._handle(); ex.stage_bam()
    ex.identify_junctions_from_BAM()
    t = np.sort(ex.junction_table(), order=["tid", "start", "end", "strand"])
    st = ex.stats()
    ex.close()
    return t, st


whole, st_whole = run(1, 0, 1)
for w in worlds:
    reads = 0; n = 0
    for r in range(w):
        th, sh = run(1, r, w)
        tf, sf = run(2, r, w)
        tr, sr = run(2, r, w, resident=True)
        assert sf["host_parse_s"] == 0.0 and sr["host_parse_s"] == 0.0, ("device feeder declined", w, r)
        assert sf["reads"] == sh["reads"] == sr["reads"], (w, r, sf["reads"], sh["reads"], sr["reads"])
        for f in F:
            assert np.array_equal(tf[f], th[f]) and np.array_equal(tr[f], th[f]), (w, r, f)
        reads += sh["reads"]; n += len(th)
    assert reads == st_whole["reads"] and n == len(whole), (w, reads, st_whole["reads"], n, len(whole))
print("ok", worlds, st_whole["reads"], len(whole))
