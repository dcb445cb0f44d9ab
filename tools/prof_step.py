"""Developer helper: times each call of the resident-batch step (used under gpurun / ncu)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import regtools_b200 as rt
from regtools_b200.distributed import _header_contigs
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
bam = bench.ensure_bam("c2", reads, 6)
ld = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=0)
t0 = time.perf_counter(); tid, pos, meta, off, cig = ld.load_batch(); print("load_batch s", time.perf_counter() - t0); ld.close()
n_nops = int(np.count_nonzero((cig & 0xF) == 3))
d = [torch.from_numpy(x.view(np.int32)).cuda() for x in (tid, pos, meta, off, cig)]
ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=0, profile=True)
ex.set_contigs(_header_contigs(bam))
stream = torch.cuda.current_stream().cuda_stream
for i in range(steps):
    torch.cuda.synchronize(); T = [time.perf_counter()]
    ex.clear(); torch.cuda.synchronize(); T.append(time.perf_counter())
    ex.scan_batch(*d, first_ordinal=0, n_junction_ops=n_nops, stream=stream); T.append(time.perf_counter())
    torch.cuda.synchronize(); T.append(time.perf_counter())
    ex.finalize(stream); T.append(time.perf_counter())
    t = ex.junction_table(); T.append(time.perf_counter())
    print("step", i, "clear %.3f scan_call %.3f scan_sync %.3f finalize %.3f table %.3f ms" % tuple(1e3 * (b - a) for a, b in zip(T, T[1:])), len(t))
print(ex.stats())
