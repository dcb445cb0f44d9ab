"""Worker of tests/test_gpu_zz_exchange.py (one process per GPU under torch.distributed.run): every rank extracts its contig
shard, rtjx_gather moves the shard tables to rank 0 over NCCL inside the library, and rank 0 compares the merged table and
its BED12 with a single-GPU run over the whole file."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import regtools_b200 as rt  # noqa: E402


def main():
    bam = sys.argv[1]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt.copy_(torch.frombuffer(bytearray(rt.JunctionsExtractor.comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(idt, 0)
    comm_id = bytes(idt.cpu().numpy().tobytes())
    for rep, inflate_mode in enumerate((0, 1, 0)):           # device feeder, host feeder, and a second handle on the same communicator
        ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local, shard_rank=rank, shard_world=world, inflate_mode=inflate_mode)
        if rep == 0:
            ex.comm_init(comm_id, rank, world)
        ex.identify_junctions_from_BAM()
        own = len(ex.junction_table())
        ex.gather(0)
        merged = ex.junction_table()
        if rank == 0:
            one = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=local)
            one.identify_junctions_from_BAM()
            want = one.junction_table()
            import io
            a, b = io.StringIO(), io.StringIO()
            one.print_all_junctions(a)
            ex.print_all_junctions(b)
            one.close()
            assert len(merged) == len(want) and len(merged) > own > 0, (len(merged), len(want), own)
            # every field the reference's Junction has; first_ord (an ordinal inside a shard's own stream) is bookkeeping
            for f in ("tid", "start", "end", "thick_start", "thick_end", "read_count", "name_index", "strand", "left_ok", "right_ok"):
                assert np.array_equal(merged[f], want[f]), f"merged table differs from the single-GPU table in {f}"
            assert a.getvalue() == b.getvalue() and len(a.getvalue()) > 1000
        else:
            assert len(merged) == own                         # the other ranks keep their shard
        ex.close()
    dist.barrier()
    if rank == 0:
        print("EXCHANGE_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
