"""Developer helper (torchrun, N >= 2): where the resident multi-GPU step spends its time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench, regtools_b200 as rt
from regtools_b200.distributed import all_gather_tables, _header_contigs
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bam = os.path.join(bench.SCRATCH, f"rtjx_c2x{world}_{10_000_000 * world}_l6.bam")
if rank == 0: bench.ensure_bam(f"c2x{world}", 10_000_000 * world, 6)
dist.barrier()
ld = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local, shard_rank=rank, shard_world=world)
arrs = ld.load_batch(); ld.close()
n_n = int(np.count_nonzero((arrs[4] & 0xF) == 3))
d = [torch.from_numpy(x.view(np.int32)).to(dev) for x in arrs]
ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local)
ex.set_contigs(_header_contigs(bam))
stream = torch.cuda.current_stream().cuda_stream
T = np.zeros(6)
for it in range(13):
    t = [time.perf_counter()]
    ex.clear(); t.append(time.perf_counter())
    ex.scan_batch(*d, n_junction_ops=n_n, stream=stream); t.append(time.perf_counter())
    ex.finalize(stream); t.append(time.perf_counter())
    tab = ex.junction_table(); t.append(time.perf_counter())
    tabs = all_gather_tables(tab, dev); t.append(time.perf_counter())
    dist.barrier(); torch.cuda.synchronize(); t.append(time.perf_counter())
    if it >= 3: T += np.diff(t)
# phase timing of the exchange itself (same steps as distributed.all_gather_tables)
from regtools_b200 import distributed as D
isz = rt.JUNCTION_DTYPE.itemsize
P = np.zeros(7)
for it in range(13):
    torch.cuda.synchronize(); p = [time.perf_counter()]
    mine_np = np.ascontiguousarray(tab, dtype=rt.JUNCTION_DTYPE).view(np.uint8).reshape(-1)
    cap = D._SLOT
    stage = D._pinned(cap, "send"); sview = stage.numpy()
    sview[:8] = np.frombuffer(np.int64(len(tab)).tobytes(), np.uint8); sview[8:8 + mine_np.size] = mine_np
    p.append(time.perf_counter())
    mine, out = D._device_slots(cap, world, dev); used = 8 + mine_np.size
    mine[:used].copy_(stage[:used], non_blocking=True)
    torch.cuda.synchronize(); p.append(time.perf_counter())
    dist.all_gather_into_tensor(out, mine)
    p.append(time.perf_counter())
    torch.cuda.synchronize(); p.append(time.perf_counter())
    host = D._pinned(world * cap, "recv"); host[:world * cap].copy_(out, non_blocking=True)
    torch.cuda.current_stream().synchronize(); p.append(time.perf_counter())
    h = host.numpy()
    sizes = [int(np.frombuffer(h[r * cap:r * cap + 8].tobytes(), np.int64)[0]) for r in range(world)]
    res = [h[r * cap + 8:r * cap + 8 + sizes[r] * isz].view(rt.JUNCTION_DTYPE).copy() for r in range(world)]
    p.append(time.perf_counter())
    dist.barrier(); torch.cuda.synchronize(); p.append(time.perf_counter())
    if it >= 3: P += np.diff(p)
print(f"\nrank {rank}: slot {D._SLOT} stage {P[0]*100:.3f} h2d {P[1]*100:.3f} nccl_call {P[2]*100:.3f} nccl_wait {P[3]*100:.3f} d2h {P[4]*100:.3f} split {P[5]*100:.3f} barrier {P[6]*100:.3f} ms", flush=True)
print(f"\nrank {rank}: clear {T[0]*100:.2f} scan_call {T[1]*100:.2f} finalize {T[2]*100:.2f} table {T[3]*100:.2f} gather {T[4]*100:.2f} barrier {T[5]*100:.2f} ms/step (avg of 10)", flush=True)
dist.destroy_process_group()
