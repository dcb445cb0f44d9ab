"""GPU parity of the warp-pipelined cigar_scan (variant 8, opt-in; kernels.cu cigar_scan_pipe_kernel) at the kernel level:
every ring / occupancy configuration, batch sizes around the 128-alignment warp tile and the persistent grid, tiles denser
than the staged slab window (ops read from global memory), alignments with many N ops (one lane per N op), and the chunked
candidate reservation (padding entries must not be counted); plus the intron-motif / variant-region / barcode modes of that
kernel against the default kernel.  Contract: the table equals the oracle's (parse_alignment_into_junctions, junctions_extractor.cc:377-497;
junction_qc :160-170; add_junction :174-235)."""
import numpy as np
import pytest

import synth
from test_gpu_parity import run_gpu_batch, run_oracle_batch, tables_equal

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


@pytest.mark.parametrize("cfg", [0, 1, 2, 3])
@pytest.mark.parametrize("n_reads", [1, 127, 128, 129, 131, 4099, 70001])
def test_ring_configs_and_ragged_sizes(cfg, n_reads):
    arrs = synth.random_batch(200 + n_reads, n_reads, spliced_frac=0.6 if n_reads < 5000 else 0.2)
    g_tab, g_bed, st = run_gpu_batch(arrs, 0, variant=8, cfg=cfg)
    o_tab, o_bed = run_oracle_batch(arrs, 0)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed
    assert st["candidates"] == synth.count_n_ops(arrs[4]) - (1 if n_reads > 10 else 0)     # padding of the chunks is not a candidate


def _long_read_batch(seed, n_reads):
    """Long-read style alignments: 1-12 N ops each, mixed with D / I / S / = / X / H / P, three contigs, some unspliced."""
    rng = np.random.default_rng(seed)
    reads = []
    for _ in range(n_reads):
        t = int(rng.integers(0, 3))
        p = int(rng.integers(0, 200000))
        ops = []
        if rng.random() < 0.3:
            ops = [int(rng.integers(50, 2000)) << 4 | 0]
        else:
            if rng.random() < 0.2:
                ops.append(int(rng.integers(1, 30)) << 4 | int(rng.choice([4, 5])))
            for _k in range(int(rng.integers(1, 13))):
                ops.append(int(rng.integers(1, 300)) << 4 | int(rng.choice([0, 0, 0, 7])))
                r = rng.random()
                if r < 0.15:
                    ops.append(int(rng.integers(1, 5)) << 4 | int(rng.choice([1, 2, 8, 6])))
                    ops.append(int(rng.integers(1, 100)) << 4 | 0)
                ops.append(int(rng.choice([60, 70, 75, 120, 1000, 500000, 500001])) << 4 | 3)
            if rng.random() < 0.9:
                ops.append(int(rng.integers(1, 300)) << 4 | 0)
        reads.append((t, p, int(rng.choice([0, 16, 99, 147])), 60, int(rng.choice([ord("+"), ord("-"), 0, ord(".")])), ops))
    reads.sort(key=lambda r: (r[0], r[1]))
    return synth.batch_from_reads(reads)


@pytest.mark.parametrize("strandness", [0, 1])
@pytest.mark.parametrize("cfg", [0, 2])
def test_many_junctions_per_alignment(cfg, strandness):
    arrs = _long_read_batch(11 + cfg, 30000)
    g_tab, g_bed, st = run_gpu_batch(arrs, strandness, variant=8, cfg=cfg, m=70)
    o_tab, o_bed = run_oracle_batch(arrs, strandness, m=70)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed
    assert st["candidates"] == synth.count_n_ops(arrs[4])


@pytest.mark.parametrize("known", [True, False])
@pytest.mark.parametrize("strandness", [0, 1, 2])
def test_random_batches_match_oracle_and_the_block_tiled_kernel(strandness, known):
    arrs = synth.random_batch(31 + strandness, 50000)
    g_tab, g_bed, st = run_gpu_batch(arrs, strandness, variant=8, known=known)
    b_tab, b_bed, _ = run_gpu_batch(arrs, strandness, variant=5, known=known)
    o_tab, o_bed = run_oracle_batch(arrs, strandness)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed == b_bed and np.array_equal(g_tab, b_tab)
    assert st["candidates"] == synth.count_n_ops(arrs[4]) - 1


def test_large_resident_batch_in_three_launches():
    arrs = tuple(synth.random_batch(9, 600000, spliced_frac=0.1, catalog_per_contig=400))
    g_tab, g_bed, st = run_gpu_batch(arrs, 0, variant=8, device_resident=True, split=3)
    o_tab, o_bed = run_oracle_batch(arrs, 0)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed
    assert st["candidates"] == synth.count_n_ops(arrs[4]) - 1


def test_special_modes_of_the_pipelined_kernel_equal_the_default_kernel(golden_dir, motif_fastas, tmp_path):
    """Intron-motif strand mode, batched variant regions and `-b` barcodes through scan_variant=8 (its per-candidate emit path)
    against the block-per-tile kernel, which the motif / regions / barcode modules pin to the oracle and the reference goldens."""
    import os
    import subprocess
    import regtools_b200 as rt
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bam = str(tmp_path / "gen.bam")
    subprocess.check_call([os.path.join(root, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "300000", "--seed", "11",
                           "--barcodes", "500"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    rng = np.random.default_rng(5)
    regions = [f"{c}:{s}-{s + w}" for c, s, w in zip(rng.choice(["1", "10", "2"], 80), rng.integers(1, 1900000, 80), rng.choice([200, 2000, 20000], 80))]
    out = {}
    for variant in (5, 8):
        ex = rt.JunctionsExtractor(bam, ".", 3, "XS", 8, 70, 500000, motif_fastas["synth"], scan_variant=variant)
        ex.identify_junctions_from_BAM()
        motif = ex.junction_table()
        ex.close()
        ex = rt.JunctionsExtractor.from_region(bam, ".", 0, "XS", 8, 70, 500000, scan_variant=variant)
        reg = ex.identify_junctions_in_regions(regions)
        ex.close()
        ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, scan_variant=variant)
        ex.output_barcodes_file_ = str(tmp_path / f"bc{variant}.tsv")
        ex.output_file_ = str(tmp_path / f"j{variant}.bed")
        ex.identify_junctions_from_BAM()
        ex.print_all_junctions()
        ex.close()
        out[variant] = (motif, reg, open(tmp_path / f"bc{variant}.tsv").read(), open(tmp_path / f"j{variant}.bed").read())
    assert len(out[5][0]) > 100 and np.array_equal(out[5][0], out[8][0])
    assert sum(len(t) for t in out[5][1]) > 100 and all(np.array_equal(a, b) for a, b in zip(out[5][1], out[8][1]))
    assert len(out[5][2]) > 1000 and out[5][2] == out[8][2] and out[5][3] == out[8][3]
