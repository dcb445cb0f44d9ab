"""Developer measurement for SURVEY §8(f)-2 / BASELINE configs[3] (cis-splice-effects: one extractor per variant region):
50k variant-like windows on the synthetic BAM, batched on the GPU (rtjx_run_regions) against the reference's 8-arg-ctor
path run region by region on a sample (oracle/_ref/regtools_ref ctor; each call re-opens BAM + index, as the reference's
loop does — plus process start-up, so the figure is an upper bound for the reference's per-region cost).
    python tools/bench_regions.py [reads] [n_regions] > gpurun_out/regions.json"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
import regtools_b200 as rt

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
n_regions = int(sys.argv[2]) if len(sys.argv) > 2 else 50_000
config = sys.argv[3] if len(sys.argv) > 3 else "c2"
bam = bench.ensure_bam(config, reads, 6)
contig_names = bench.config_contigs(config)
ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000)
ex.identify_junctions_from_BAM()
tab = ex.junction_table()
ex.close()
rng = np.random.default_rng(1234)
pick = tab[rng.integers(0, len(tab), n_regions)]
side = rng.integers(0, 2, n_regions)
centre = np.where(side == 0, pick["start"], pick["end"]).astype(np.int64) + rng.integers(-10, 11, n_regions)
regions = [f"{contig_names[int(t)]}:{max(1, c - 500)}-{c + 500}" for t, c in zip(pick["tid"], centre)]
windows = [(max(1, int(c) - 500), int(c) + 500) for c in centre]
times = []
for rep in range(3):
    t0 = time.perf_counter()
    e = rt.JunctionsExtractor.from_region(bam, ".", 0, "XS", 8, 70, 500000)
    out = e.identify_junctions_in_regions(regions)
    t_tables = time.perf_counter() - t0
    uniq, first, variants = e.unique_junctions_in_windows(regions, windows) if rep == 2 else (None, None, None)
    st = e.stats()
    e.close()
    times.append(t_tables)
n_j = int(sum(len(t) for t in out))
ref_bin = os.path.join(ROOT, "oracle", "_ref", "regtools_ref")
sample = list(range(0, n_regions, max(1, n_regions // 40)))[:40]
t0 = time.perf_counter()
same = True
for i in sample:
    p = subprocess.run([ref_bin, "ctor", bam, regions[i], "0", "XS", "8", "70", "500000"], capture_output=True, text=True)
    lines = ["\t".join(map(str, [contig_names[int(j["tid"])], j["thick_start"], j["thick_end"], "JUNC%08d" % j["name_index"], j["read_count"], chr(j["strand"]),
                                 j["start"], j["end"], int(j["left_ok"]), int(j["right_ok"])])) + "\n" for j in out[i]]
    same = same and p.returncode == 0 and p.stdout == "".join(lines)
ref_per_region = (time.perf_counter() - t0) / len(sample)
best = min(times)
print(json.dumps({
    "workload": f"{n_regions} windows of 1 kb around junction ends of the synthetic {reads}-read {config} BAM (BASELINE configs[3] shape, variants +-10 bp from donors/acceptors)",
    "unique_junction_set": None if uniq is None else {"unique_junctions": int(len(uniq)), "junction_variant_pairs": int(sum(len(v) for v in variants))},
    "batched_s": best, "batched_runs_s": times, "regions_per_s": n_regions / best, "junction_rows": n_j,
    "kernel_launches": int(st["kernel_launches"]), "reads_streamed": int(st["reads"]),
    "reference_s_per_region": ref_per_region, "reference_sample": len(sample), "reference_extrapolated_s": ref_per_region * n_regions,
    "speedup_vs_reference_extrapolated": ref_per_region * n_regions / best, "sample_outputs_identical_to_reference": bool(same),
    "note": "reference = oracle/_ref/regtools_ref ctor <bam> <region> ... per region (process start + BAM/index open + iterate); single thread"}))
