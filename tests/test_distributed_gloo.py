"""CPU, world_size 2, gloo: the N>1 path's host logic — contig shard plan, all-gatherv of junction
tables, merge + renaming + BED12 on rank 0.  The per-rank extraction (CUDA on the GPU box) is stood in
for by the oracle restricted to the rank's contigs; everything else is the product code."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYN = os.path.join(ROOT, "tests", "golden", "kat", "synth.bam")


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import regtools_b200 as rt
    from regtools_b200.distributed import all_gather_tables, merge_tables
    from oracle_py import Oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    assign = rt.plan_shards(SYN, world)
    names = ["1", "10", "2"]
    parts = []
    for tid, r in enumerate(assign):
        if r != rank:
            continue
        o = Oracle(8, 70, 500000, 0)
        o.extract_bam(SYN, names[tid])
        t = o.table()
        part = np.zeros(len(t), rt.JUNCTION_DTYPE)
        for f in ("tid", "start", "end", "thick_start", "thick_end", "read_count", "strand", "left_ok", "right_ok"):
            part[f] = t[f]
        part["first_ord"] = t["name_index"]
        parts.append(part)
    mine = np.concatenate(parts) if parts else np.zeros(0, rt.JUNCTION_DTYPE)
    tables = all_gather_tables(mine)
    assert sum(len(t) for t in tables) >= len(mine)
    if rank == 0:
        m = merge_tables(SYN, tables)
        m.output_file_ = out_path
        m.print_all_junctions()
        m.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_extract_matches_whole_file(tmp_path):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_py import Oracle
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = str(tmp_path / "merged.bed")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    whole = Oracle(8, 70, 500000, 0)
    whole.extract_bam(SYN)
    assert open(out).read() == whole.bed12()
