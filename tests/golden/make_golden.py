#!/usr/bin/env python
"""Generates tests/golden/kat/: small BAMs that exercise what the reference's own fixtures do not
(S/H/P/=/X/B ops, adjacent and trailing N, odd XS tags, multi-contig naming, anchor OR across reads,
intron-length bounds, unmapped-flagged reads, records straddling BGZF blocks, region queries) and the
outputs of the UNMODIFIED reference (oracle/_ref/regtools_ref, built from /root/reference by
oracle/Makefile) on them.  Run in the dev container only; the results are committed.

    python tests/golden/make_golden.py
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bamio  # noqa: E402

REF = os.path.join(ROOT, "oracle", "_ref", "regtools_ref")
BAMGEN = os.path.join(ROOT, "tools", "bamgen")
OUT = os.path.join(HERE, "kat")

XSP, XSM = b"XSA+", b"XSA-"
CONTIGS = [("1", 30_000_000), ("10", 20_000_000), ("2", 25_000_000)]

# (tid, pos, cigar, flag, mapq, aux) — SURVEY.md §8a KAT table and more
KAT = [
    (0, 1000, "50M100N50M", 0, 60, XSP), (0, 1000, "10S40M100N50M", 0, 0, XSP),
    (0, 1000, "5H50M100N50M5H", 0, 60, XSP), (0, 1000, "20M2X28M100N50M", 0, 60, XSP),
    (0, 1000, "20=30M100N25=25M", 0, 60, XSP), (0, 1000, "50M100N20M1I29M", 0, 60, XSP),
    (0, 1000, "50M100N20M3D30M", 0, 60, XSP),
    (0, 2000, "50M100N200N50M", 0, 60, XSP), (0, 3000, "50M100N30M200N20M", 16, 60, XSM),
    (0, 4000, "50M100N", 0, 60, XSP), (0, 5000, "50M69N50M", 0, 60, XSP), (0, 5500, "50M70N50M", 0, 60, XSP),
    (0, 6000, "7M100N50M", 0, 60, XSP), (0, 6000, "50M100N7M", 0, 60, XSP),
    (0, 6957, "50M100N7M", 0, 60, XSP), (0, 7000, "7M100N50M", 0, 60, XSP),
    (0, 8000, "50M100N50M", 0, 60, b"XSA."), (0, 8000, "50M100N50M", 0, 60, b""),
    (0, 8500, "50M100N50M", 0, 60, b"XSC\x2b"), (0, 8600, "50M100N50M", 0, 60, b"NHC\x01XSA-"),
    (0, 8700, "50M100N50M", 0, 60, b"XSZabc\0XSA+"), (0, 8800, "50M100N50M", 0, 60, b"ZZBc\x03\0\0\0\x01\x02\x03XSA+"),
    (0, 9500, "50M3P100N50M", 0, 60, XSP), (0, 10000, "50M100N50M", 4, 0, XSP),
    (0, 10000, "50M100N50M", 256 | 512 | 1024 | 2048, 0, XSP),
    (0, 11000, "50M500001N50M", 0, 60, XSP), (0, 11000, "50M500000N50M", 0, 60, XSP),
    (0, 12000, "30M5S", 0, 60, b""), (0, 12000, "100M", 0, 60, b""), (0, 12100, "", 4, 0, b""),
    (0, 13000, "10M80N10M80N10M80N10M80N10M80N10M80N10M80N10M80N10M", 99, 60, XSM),
    (0, 14000, "25M100N25M", 99, 60, b""), (0, 14000, "25M100N25M", 147, 60, b""),
    (0, 14000, "25M100N25M", 83, 60, b""), (0, 14000, "25M100N25M", 163, 60, b""),
    (0, 14000, "25M100N25M", 65, 60, b""), (0, 14000, "25M100N25M", 177, 60, b""),
    (1, 100, "50M100N50M", 0, 60, XSP), (1, 200, "50M100N50M", 0, 60, XSM),
    (2, 100, "50M100N50M", 0, 60, XSP), (2, 5_000_000, "40M3000N60M", 16, 3, XSM),
]


def build():
    os.makedirs(OUT, exist_ok=True)
    # kat.bam: tiny BGZF blocks so records straddle block boundaries
    recs = [bamio.record(t, p, c, f, q, a, name=b"k%03d" % i) for i, (t, p, c, f, q, a) in enumerate(KAT)]
    bam = os.path.join(OUT, "kat.bam")
    bamio.write_bam(bam, CONTIGS, recs, block_size=300)
    subprocess.check_call([BAMGEN, "index", bam])
    # synth.bam: generator output (three contigs, ~6k reads), normal block size
    syn = os.path.join(OUT, "synth.bam")
    subprocess.check_call([BAMGEN, "gen", "--out", syn, "--config", "tiny", "--reads", "6000", "--seed", "99", "--threads", "2"],
                          stdout=subprocess.DEVNULL)
    runs = {
        "kat": [["-s", "XS"], ["-s", "XS", "-a", "0"], ["-s", "RF"], ["-s", "FR", "-a", "0"], ["-s", "XS", "-t", "NH"],
                ["-s", "XS", "-r", "1:5000-6200"], ["-s", "XS", "-r", "10"], ["-s", "XS", "-r", "2:5,000,050-5,000,060", "-a", "0"],
                ["-s", "XS", "-m", "0", "-a", "0", "-M", "4294967295"]],
        "synth": [["-s", "XS"], ["-s", "RF", "-a", "20"], ["-s", "FR"], ["-s", "XS", "-r", "10:500000-900000"],
                  ["-s", "XS", "-r", "2"], ["-s", "XS", "-m", "100", "-M", "5000", "-a", "1"]],
    }
    manifest = []
    for stem, arglists in runs.items():
        for i, args in enumerate(arglists):
            out = os.path.join(OUT, f"{stem}.{i}.bed")
            p = subprocess.run([REF, "junctions", "extract"] + args + [os.path.join(OUT, stem + ".bam")],
                               capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
            open(out, "w").write(p.stdout)
            manifest.append(f"{stem}.bam\t{stem}.{i}.bed\t{' '.join(args)}")
    # second caller: the 8-arg ctor (min_intron := min_anchor quirk, unfiltered get_all_junctions)
    for i, (region, anchor) in enumerate([("1:5000-6200", 8), ("1:900-1300", 8), ("2", 8)]):
        out = os.path.join(OUT, f"kat.ctor{i}.tsv")
        p = subprocess.run([REF, "ctor", os.path.join(OUT, "kat.bam"), region, "0", "XS", str(anchor), "70", "500000"],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        open(out, "w").write(p.stdout)
        manifest.append(f"kat.bam\tkat.ctor{i}.tsv\tctor {region} 0 XS {anchor} 70 500000")
    open(os.path.join(OUT, "MANIFEST.tsv"), "w").write("\n".join(manifest) + "\n")
    print(f"wrote {len(manifest)} golden outputs to {OUT}")


MOTIF_RUNS = [
    ("kat.bam", "kat", ["-s", "XS"]), ("kat.bam", "kat", ["-s", "intron-motif", "-a", "0"]),
    ("kat.bam", "kat", ["-s", "RF", "-m", "0", "-a", "0", "-M", "4294967295"]),
    ("synth.bam", "synth", ["-s", "XS"]), ("synth.bam", "synth", ["-s", "RF", "-a", "20"]),
    ("synth.bam", "synth", ["-s", "intron-motif"]), ("synth.bam", "synth", ["-s", "FR", "-r", "2"]),
    ("kat.bam", "kat_no10", ["-s", "XS"]),
]


def build_motif():
    """tests/golden/motif/: the reference with a FASTA as second positional argument (strand from the intron motif,
    junctions_extractor.cc:325-359) on the FASTAs tests/fasta_fixture.py generates."""
    import tempfile
    import fasta_fixture as ff
    out_dir = os.path.join(HERE, "motif")
    os.makedirs(out_dir, exist_ok=True)
    manifest = []
    with tempfile.TemporaryDirectory() as tmp:
        fas = {"kat": ff.write_kat_fasta(os.path.join(tmp, "kat.fa")), "synth": ff.write_synth_fasta(os.path.join(tmp, "synth.fa")),
               "kat_no10": ff.write_kat_fasta(os.path.join(tmp, "kat_no10.fa"), drop="10")}
        for i, (bam, fa, args) in enumerate(MOTIF_RUNS):
            p = subprocess.run([REF, "junctions", "extract"] + args + [os.path.join(OUT, bam), fas[fa]], capture_output=True, text=True)
            name = f"motif.{i}.bed"
            open(os.path.join(out_dir, name), "w").write(p.stdout)
            err = [l[l.index("Unable"):] for l in p.stderr.splitlines() if "Unable" in l]      # "Unknown cigar P" has no newline
            manifest.append(f"{bam}\t{fa}\t{name}\t{p.returncode}\t{' '.join(args)}\t{err[0] if err else ''}")
    open(os.path.join(out_dir, "MANIFEST.tsv"), "w").write("\n".join(manifest) + "\n")
    print(f"wrote {len(manifest)} intron-motif golden outputs to {out_dir}")


BC2_VARIANTS = {
    "xs": ["-s", "XS"],
    "xs_a0": ["-s", "XS", "-a", "0"],
    "rf_m50": ["-s", "RF", "-m", "50"],
    "fr_region": ["-s", "FR", "-r", "10:1-60000"],
}


def build_barcodes():
    """tests/golden/barcodes/: `-b` single-cell mode of the reference (set_junction_barcode, print_barcodes:
    junctions_extractor.cc:362-374, .h:99-111) on a fixture with CB:Z tags (5 and 60 distinct barcodes per locus, reads
    without CB, CB before and after XS)."""
    import random
    out_dir = os.path.join(HERE, "barcodes")
    os.makedirs(out_dir, exist_ok=True)
    rnd = random.Random(7)
    bcs = [("".join(rnd.choice("ACGT") for _ in range(16)) + "-1").encode() for _ in range(60)]
    loci = [(1000, "50M100N50M"), (5000, "40M300N60M"), (9000, "30M1000N70M"), (9000, "30M1000N20M500N50M")]
    reads = []
    for _ in range(400):
        pos, cg = rnd.choice(loci)
        r = rnd.random()
        if r < 0.05:
            aux = b"XSA+"
        else:
            bc = rnd.choice(bcs[:5 if pos == 1000 else 60])
            aux = b"XSA+CBZ" + bc + b"\0" if r < 0.9 else b"CBZ" + bc + b"\0XSA-"
        reads.append((pos, cg, aux))
    reads.sort(key=lambda x: x[0])
    recs = [bamio.record(0, p, c, 0, 60, a, name=b"q%04d" % i) for i, (p, c, a) in enumerate(reads)]
    bam = os.path.join(out_dir, "bc.bam")
    bamio.write_bam(bam, [("1", 100000)], recs)
    subprocess.check_call([BAMGEN, "index", bam])
    p = subprocess.run([REF, "junctions", "extract", "-s", "XS", "-b", os.path.join(out_dir, "bc.barcodes"), "-o",
                        os.path.join(out_dir, "bc.bed"), bam], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    # bc2: the adversarial fixture of tests/bc_fixture.py (3 contigs, a hot junction with hundreds of barcodes, proxy-2
    # strands, QC- and anchor-filtered junctions, CB behind other tags), several flag sets, all by the unmodified reference
    import bc_fixture
    bam2 = bc_fixture.make_barcode_bam(os.path.join(out_dir, "bc2.bam"), seed=11, n_reads=3000)
    for tag, args in BC2_VARIANTS.items():
        p = subprocess.run([REF, "junctions", "extract"] + args + ["-b", os.path.join(out_dir, f"bc2.{tag}.barcodes"), "-o",
                            os.path.join(out_dir, f"bc2.{tag}.bed"), bam2], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        with open(os.path.join(out_dir, f"bc2.{tag}.warnings"), "w") as f:
            f.write(str(p.stderr.count("WARNING: No CB tag found for alignment (id = 0)")) + "\n")
    print(f"wrote the barcode golden to {out_dir}")


REF_ANN = os.path.join(ROOT, "oracle", "_ref", "regtools_ref_annotate")
ANN_CASES = {"s1": dict(seed=1), "s2": dict(seed=2, header=True), "s3": dict(seed=3, crlf=True)}


def build_annotate():
    """tests/golden/annotate/: `junctions annotate`.  hcc1395.* = the reference's own integration-test
    inputs and expected output (tests/integration-test/data/{bed,fa,gtf,junctions-annotate}); s<N>.* = generated cases
    (tests/ann_fixture.py) annotated by the UNMODIFIED reference (oracle/_ref/regtools_ref_annotate), with and without -S.
    The FASTA of the generated cases is rebuilt by ann_fixture.write_fasta (explicit LCG), not committed."""
    import shutil
    import ann_fixture
    out_dir = os.path.join(HERE, "annotate")
    os.makedirs(out_dir, exist_ok=True)
    d = "/root/reference/tests/integration-test/data"
    for src, dst in (("bed/test_hcc1395_junctions.bed", "hcc1395.bed"), ("fa/test_chr22.fa", "hcc1395.fa"),
                     ("gtf/test_ensemble_chr22.gtf", "hcc1395.gtf"), ("junctions-annotate/expected-annotate.out", "hcc1395.expected.tsv")):
        shutil.copyfile(os.path.join(d, src), os.path.join(out_dir, dst))
        os.chmod(os.path.join(out_dir, dst), 0o644)
    fa = ann_fixture.write_fasta(os.path.join("/tmp", "ann_golden_ref.fa"))
    for tag, kw in ANN_CASES.items():
        tmp = os.path.join("/tmp", "ann_golden_" + tag)
        bed, _, gtf = ann_fixture.make_annotation_case(tmp, **kw)
        shutil.copyfile(bed, os.path.join(out_dir, tag + ".bed"))
        shutil.copyfile(gtf, os.path.join(out_dir, tag + ".gtf"))
        for flag, suffix in (([], ""), (["-S"], ".S")):
            p = subprocess.run([REF_ANN, "junctions", "annotate"] + flag + ["-o", os.path.join(out_dir, f"{tag}{suffix}.expected.tsv"), bed, fa, gtf],
                               capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
    print(f"wrote the annotate goldens to {out_dir}")


if __name__ == "__main__":
    if "--annotate-only" in sys.argv:
        build_annotate()
        sys.exit(0)
    if "--barcodes-only" in sys.argv:
        build_barcodes()
        sys.exit(0)
    if not os.path.exists(REF):
        sys.exit("oracle/_ref/regtools_ref missing: run `make -C oracle ref` in the dev container first")
    if "--motif-only" not in sys.argv:
        build()
        build_barcodes()
        build_annotate()
    build_motif()
