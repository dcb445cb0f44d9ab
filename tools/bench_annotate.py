#!/usr/bin/env python
"""Measurement of `junctions annotate` — a side bench, NOT bench.py's contract line (that one stays on the
junctions-extract hot path).  Prints one JSON line:

  value          junctions/s of the B200 path through the public call (JunctionsAnnotator.annotate_all: GTF + BED + FASTA
                 files in, TSV file out; parsing, H2D, kernel, D2H and the text writer all inside the timed region)
  cpu_baseline   the UNMODIFIED reference (oracle/_ref/regtools_ref_annotate, single thread — it has no threading) or, where
                 that binary is absent, the oracle port, on a bounded sample of the same junctions, same GTF and FASTA
  parity         the B200 output of the sample lines compared byte for byte with the CPU arm's

Workload: tests/ann_fixture.py gene models on `--contigs` contigs of `--mb` Mb (about 35 transcripts per Mb), every junction
line repeated until `--junctions` lines.  `--impl emul` times the host emulation harness instead of the GPU (a check of
this script on a box without a GPU; never a result).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ann_fixture  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--contigs", type=int, default=8)
    ap.add_argument("--mb", type=int, default=30)
    ap.add_argument("--junctions", type=int, default=2_000_000)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--sample", type=int, default=200_000, help="junction lines given to the CPU arm")
    ap.add_argument("--impl", default="ours", choices=["ours", "emul"])
    ap.add_argument("--dir", default="/tmp/rtjx_bench_annotate")
    a = ap.parse_args()

    contigs = [("chr%d" % (i + 1), a.mb * 1_000_000) for i in range(a.contigs)] + [("chrEmpty", 1_000_000)]
    d = os.path.join(a.dir, f"c{a.contigs}_m{a.mb}")
    os.makedirs(d, exist_ok=True)
    bed0, fa, gtf = os.path.join(d, "junctions.bed"), os.path.join(d, "ref.fa"), os.path.join(d, "ann.gtf")
    if not (os.path.exists(bed0) and os.path.exists(gtf)):
        ann_fixture.make_annotation_case(d, 1234, contigs=contigs, fasta=False)
    if not os.path.exists(fa):
        ann_fixture.write_fasta_numpy(fa, contigs)
    lines = open(bed0).read().splitlines(keepends=True)
    reps = max(1, -(-a.junctions // len(lines)))
    bed = os.path.join(d, f"junctions_x{reps}.bed")
    if not os.path.exists(bed):
        with open(bed, "w") as f:
            for _ in range(reps):
                f.writelines(lines)
    n = reps * len(lines)
    sample = os.path.join(d, f"sample_{min(a.sample, n)}.bed")
    with open(sample, "w") as f:
        f.writelines((lines * reps)[:min(a.sample, n)])
    n_sample = min(a.sample, n)

    out = os.path.join(d, "ours.tsv")
    if a.impl == "ours":
        import regtools_b200 as rt

        def step(b, o):
            an = rt.JunctionsAnnotator(b, fa, gtf)
            an.output_file_ = o
            return an.annotate_all()
    else:
        exe = os.path.join(ROOT, "build", "emul", "annotate_emul")
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "-s"])

        def step(b, o):
            subprocess.check_call([exe, "-o", o, b, fa, gtf], stderr=subprocess.DEVNULL)
            return n
    for _ in range(a.warmup):
        step(bed, out)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step(bed, out)
    dt = (time.perf_counter() - t0) / a.steps

    # CPU arm on the sample + parity of the same lines
    ref = os.path.join(ROOT, "oracle", "_ref", "regtools_ref_annotate")
    kind = "reference"
    if os.path.exists(ref):
        cmd = [ref, "junctions", "annotate", "-o", os.path.join(d, "cpu.tsv"), sample, fa, gtf]
    else:
        kind = "port"
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
        cmd = [os.path.join(ROOT, "oracle", "_ref", "ja_oracle"), "-o", os.path.join(d, "cpu.tsv"), sample, fa, gtf]
    t0 = time.perf_counter()
    subprocess.check_call(cmd, stderr=subprocess.DEVNULL)
    cpu_dt = time.perf_counter() - t0
    step(sample, os.path.join(d, "ours_sample.tsv"))
    parity = open(os.path.join(d, "ours_sample.tsv"), "rb").read() == open(os.path.join(d, "cpu.tsv"), "rb").read()

    n_tx = sum(1 for _ in {l.split('transcript_id "')[1].split('"')[0] for l in open(gtf) if "\texon\t" in l})
    print(json.dumps({
        "metric": "junctions/sec through junctions annotate (files in, TSV out)", "impl": a.impl, "value": n / dt, "unit": "junctions/s",
        "ms_per_step": dt * 1e3, "steps": a.steps, "warmup": a.warmup, "higher_is_better": True, "data": "synthetic",
        "config": {"workload": f"{n} BED12 junctions x GTF of {n_tx} transcripts on {a.contigs} x {a.mb} Mb + FASTA", "repeats_of_distinct_lines": reps},
        "cpu_baseline": {"value": n_sample / cpu_dt, "unit": "junctions/s", "cores": 1, "kind": kind, "sample": f"first {n_sample} junction lines, same GTF and FASTA"},
        "parity": "byte-identical on the sample" if parity else "MISMATCH on the sample",
    }))
    return 0 if parity else 1


if __name__ == "__main__":
    sys.exit(main())
