"""GPU parity of cigar_scan variant 7 (gather variant, kernels.cu cigar_scan_gather_kernel): an opt-in A/B build written at
the end of round 1 for the occupancy experiment described in DESIGN.md §9 — it stages only cig_off + the CIGAR slab and
gathers pos/meta/tid per work item.  Same contract as the default scan: every table must equal the oracle's
(parse_alignment_into_junctions, junctions_extractor.cc:377-497).  Sorts late in the suite: first GPU run at round end."""
import numpy as np
import pytest

import synth
from test_gpu_parity import run_gpu_batch, run_oracle_batch, tables_equal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n_reads", [1, 511, 512, 513, 4099, 70001])
def test_tile_configs_and_ragged_sizes(cfg, n_reads):
    """Batch sizes around the tile boundaries; dense splicing makes tiles overflow the staged slab window and the 96-entry
    candidate stage (several rounds, ops read from global memory)."""
    arrs = synth.random_batch(100 + n_reads, n_reads, spliced_frac=0.6 if n_reads < 5000 else 0.2)
    g_tab, g_bed, _ = run_gpu_batch(arrs, 0, variant=7, cfg=cfg)
    o_tab, o_bed = run_oracle_batch(arrs, 0)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed


@pytest.mark.parametrize("strandness", [0, 1, 2])
@pytest.mark.parametrize("known", [True, False])
def test_random_batches_match_oracle(strandness, known):
    arrs = synth.random_batch(7 + strandness, 20000)
    g_tab, g_bed, st = run_gpu_batch(arrs, strandness, variant=7, known=known)
    o_tab, o_bed = run_oracle_batch(arrs, strandness)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed
    assert st["candidates"] == synth.count_n_ops(arrs[4]) - 1      # read 3 has tid -1: its N op is never emitted


def test_large_resident_batch_equals_the_default_variant():
    arrs = tuple(synth.random_batch(9, 300000, spliced_frac=0.1, catalog_per_contig=400))
    g_tab, g_bed, _ = run_gpu_batch(arrs, 0, variant=7, device_resident=True, split=3)
    d_tab, d_bed, _ = run_gpu_batch(arrs, 0, variant=5, device_resident=True, split=3)
    o_tab, o_bed = run_oracle_batch(arrs, 0)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed == d_bed and np.array_equal(g_tab, d_tab)
