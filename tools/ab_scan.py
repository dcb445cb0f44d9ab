"""Developer A/B harness for cigar_scan (run under gpurun): times the resident 10M-read C2 batch through
every scan variant / tile configuration / debug switch in ONE process and checks that the tables agree.
    python tools/ab_scan.py [reads] [steps] > gpurun_out/ab_scan.json"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import regtools_b200 as rt
from regtools_b200.distributed import _header_contigs
import bench

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
config = sys.argv[3] if len(sys.argv) > 3 else "c2"
bam = bench.ensure_bam(config, reads, 6)
ld = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=0)
tid, pos, meta, off, cig = ld.load_batch(); ld.close()
R, C = len(tid), len(cig)
n_nops = int(np.count_nonzero((cig & 0xF) == 3))
d = [torch.from_numpy(x.view(np.int32)).cuda() for x in (tid, pos, meta, off, cig)]
stream = torch.cuda.current_stream().cuda_stream
contigs = _header_contigs(bam)
alg = 16.0 * R + 4.0 * C
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
out = {"reads": R, "cigar_ops": C, "n_ops_N": n_nops, "algorithmic_bytes": alg, "peak_gbs": peak, "runs": []}
ref = None
CASES = [(5, 0, 0), (8, 0, 0), (8, 1, 0), (8, 2, 0), (8, 3, 0), (5, 0, 0)]
if os.environ.get('AB_CASES'):
    CASES = [tuple(int(x) for x in c.split(':')) for c in os.environ['AB_CASES'].split(',')]
for variant, cfg, dbg in CASES:
    ex = rt.JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=0, profile=True, scan_variant=variant, scan_cfg=cfg)
    ex.set_contigs(contigs)
    for i in range(3 + steps):
        if i == 3:
            ex.reset_stats()
        ex.clear()
        ex.scan_batch(*d, first_ordinal=0, n_junction_ops=n_nops, stream=stream)
        torch.cuda.synchronize()
    st = ex.stats()
    row = {"variant": variant, "cfg": cfg, "debug": dbg, "scan_us": 1e3 * st["scan_ms"] / st["batches"],
           "merge_us": 1e3 * st["merge_ms"] / st["batches"]}
    row["frac_of_peak"] = alg / (row["scan_us"] * 1e-6) / 1e9 / peak
    if dbg == 0:
        t = ex.junction_table()
        sig = (len(t), int(t["read_count"].sum()), int(t["name_index"].astype(np.int64).sum()), int(t["thick_start"].astype(np.int64).sum()),
               int(t["thick_end"].astype(np.int64).sum()), int(t["strand"].astype(np.int64).sum()), int(t["left_ok"].sum()), int(t["right_ok"].sum()))
        row["table_sig"] = sig
        if ref is None:
            ref = t
        row["equal_to_first"] = bool(len(t) == len(ref) and all(np.array_equal(t[f], ref[f]) for f in t.dtype.names))
    ex.close()
    out["runs"].append(row)
    print(json.dumps(row), file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
