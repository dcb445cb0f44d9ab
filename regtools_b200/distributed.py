"""Multi-GPU `junctions extract`: one process per GPU, contigs sharded across ranks, one exchange.

The junction key contains the contig, so per-rank tables are disjoint and the only collective on the
path is the final all-gather(v) of the compacted junction tables.  torch.distributed
is plumbing: NCCL (CUDA tensors over NVLink/NVSwitch) on the GPU box, gloo (CPU tensors) in the
world_size-2 CPU tests.  Names are re-ranked by (tid, first_ord) on the gathering rank, which equals
BAM order for a coordinate-sorted file.
"""
import os
from typing import List, Optional

import numpy as np

from .extractor import JUNCTION_DTYPE, JunctionsExtractor


def all_gather_tables(table: np.ndarray, device=None) -> List[np.ndarray]:
    """all-gatherv of rtjx_junction tables: sizes first, then bytes padded to the largest table."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    backend = dist.get_backend()
    dev = device if device is not None else ("cuda" if backend == "nccl" else "cpu")
    n = torch.tensor([len(table)], dtype=torch.int64, device=dev)
    sizes = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1) * JUNCTION_DTYPE.itemsize
    raw = np.zeros(cap, np.uint8)
    raw[:len(table) * JUNCTION_DTYPE.itemsize] = np.ascontiguousarray(table, dtype=JUNCTION_DTYPE).view(np.uint8)
    mine = torch.from_numpy(raw).to(dev)
    bufs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(bufs, mine)
    out = []
    for r in range(world):
        b = bufs[r].cpu().numpy()[:sizes[r] * JUNCTION_DTYPE.itemsize]
        out.append(b.view(JUNCTION_DTYPE).copy())
    return out


def merge_tables(bam: str, tables: List[np.ndarray], min_anchor: int = 8) -> JunctionsExtractor:
    """Host-only handle holding the union of per-shard tables (names re-ranked, sorted)."""
    ex = JunctionsExtractor(bam, ".", 0, "XS", min_anchor, 70, 500000, device=-1)
    ex.set_contigs(_header_contigs(bam))
    for t in tables:
        if len(t):
            ex.import_table(t)
    return ex


def _header_contigs(bam: str) -> List[str]:
    """Contig names from the BAM header (minimal reader; the header sits in the first BGZF blocks)."""
    import struct
    import zlib
    data = b""
    with open(bam, "rb") as f:
        raw = f.read(1 << 16)
        off = 0
        names: Optional[List[str]] = None
        while names is None:
            while len(raw) - off < 18:
                more = f.read(1 << 16)
                if not more:
                    raise RuntimeError("truncated BAM header")
                raw += more
            bsize = struct.unpack_from("<H", raw, off + 16)[0] + 1
            while len(raw) - off < bsize:
                more = f.read(1 << 16)
                if not more:
                    raise RuntimeError("truncated BAM header")
                raw += more
            data += zlib.decompress(raw[off + 18:off + bsize - 8], -15)
            off += bsize
            names = _try_parse_header(data)
    return names


def _try_parse_header(data: bytes):
    import struct
    if len(data) < 12:
        return None
    l_text = struct.unpack_from("<i", data, 4)[0]
    p = 8 + l_text
    if len(data) < p + 4:
        return None
    n_ref = struct.unpack_from("<i", data, p)[0]
    p += 4
    names = []
    for _ in range(n_ref):
        if len(data) < p + 4:
            return None
        l_name = struct.unpack_from("<i", data, p)[0]
        if len(data) < p + 4 + l_name + 4:
            return None
        names.append(data[p + 4:p + 4 + l_name - 1].decode())
        p += 8 + l_name
    return names


def extract_sharded(bam: str, strandness: int = 0, strand_tag: str = "XS", min_anchor: int = 8, min_intron: int = 70,
                    max_intron: int = 500000, output: Optional[str] = None, n_threads: int = 0, device: Optional[int] = None):
    """`regtools junctions extract` over torch.distributed: every rank extracts its contigs on its GPU,
    tables are all-gathered, rank 0 writes the BED12.  Returns rank 0's merged handle (else None)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    ex = JunctionsExtractor(bam, ".", strandness, strand_tag, min_anchor, min_intron, max_intron, device=device,
                            n_threads=n_threads, shard_rank=rank, shard_world=world)
    ex.identify_junctions_from_BAM()
    table = ex.junction_table()
    ex.close()
    tables = all_gather_tables(table)
    if rank != 0:
        return None
    merged = merge_tables(bam, tables, min_anchor)
    if output:
        merged.output_file_ = output
        merged.print_all_junctions()
    return merged
