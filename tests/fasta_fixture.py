"""Deterministic FASTA fixtures for the intron-motif strand mode (second positional argument of
`regtools junctions extract`, junctions_extractor.cc:325-359,548-584).

The FASTAs are generated, not committed (synth.fa is 7.5 MB): a fixed xorshift stream gives the background, and
splice-site 2-mers are planted at the junctions listed in committed reference outputs, so the files are
byte-identical wherever they are built.  tests/golden/motif/*.bed are the outputs of the UNMODIFIED reference on
them (tests/golden/make_golden.py)."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
KAT_DIR = os.path.join(HERE, "golden", "kat")

MOTIFS = [(b"GT", b"AG"), (b"GC", b"AG"), (b"AT", b"AC"), (b"CT", b"AC"), (b"CT", b"GC"), (b"GT", b"AT"), (b"gt", b"ag"), None]


def _background(n, seed):
    """n bases from a 64-bit xorshift* stream (pure integer ops: identical on every numpy version)."""
    x = np.arange(1, n + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)
    x ^= x >> np.uint64(12)
    x ^= x << np.uint64(25)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x2545F4914F6CDD1D)
    return np.frombuffer(b"ACGT", dtype=np.uint8)[((x >> np.uint64(33)) & np.uint64(3)).astype(np.int64)].copy()


def _junctions_from_bed(path):
    out = []
    for line in open(path):
        f = line.rstrip("\n").split("\t")
        left, right = (int(v) for v in f[10].split(","))
        out.append((f[0], int(f[1]) + left, int(f[2]) - right))
    return out


def _plant(seqs, junctions):
    for chrom, start, end in junctions:
        if chrom not in seqs:
            continue
        m = MOTIFS[(start * 7 + end * 13) % len(MOTIFS)]
        s = seqs[chrom]
        if m is None or end > len(s) or end - start < 4:
            continue
        s[start:start + 2] = np.frombuffer(m[0], dtype=np.uint8)
        s[end - 2:end] = np.frombuffer(m[1], dtype=np.uint8)


def _write(path, seqs, order, width=60):
    with open(path, "wb") as f:
        for name in order:
            s = seqs[name].tobytes()
            f.write(b">" + name.encode() + b" synthetic\n")
            for i in range(0, len(s), width):
                f.write(s[i:i + width] + b"\n")
    return path


def write_kat_fasta(path, drop=None):
    """FASTA for tests/golden/kat/kat.bam: contig 1 covers every KAT junction, contig 10 is 1 kb, contig 2 is so short
    that its junctions are clipped to an empty fetch.  `drop` removes one contig (the reference then throws)."""
    seqs = {"1": _background(20000, 11), "10": _background(1000, 12), "2": _background(151, 13)}
    seqs["1"][5000:5200] = np.frombuffer(b"N", dtype=np.uint8)[0]
    _plant(seqs, _junctions_from_bed(os.path.join(KAT_DIR, "kat.8.bed")))
    # the read at 3000 (50M100N30M200N20M, XS -): both introns CT..AC, so the second one is seen reverse-complemented
    for start, end in ((3050, 3150), (3180, 3380)):
        seqs["1"][start:start + 2] = np.frombuffer(b"CT", dtype=np.uint8)
        seqs["1"][end - 2:end] = np.frombuffer(b"AC", dtype=np.uint8)
    order = [c for c in ("1", "10", "2") if c != drop]
    return _write(path, seqs, order)


def write_synth_fasta(path):
    """FASTA for tests/golden/kat/synth.bam (bamgen `tiny`: contigs 1/10/2 of 3.0/2.0/2.5 Mb)."""
    seqs = {"1": _background(3000000, 21), "10": _background(2000000, 22), "2": _background(2500000, 23)}
    j = _junctions_from_bed(os.path.join(KAT_DIR, "synth.0.bed")) + _junctions_from_bed(os.path.join(KAT_DIR, "synth.2.bed"))
    _plant(seqs, j)
    return _write(path, seqs, ["1", "10", "2"], width=70)
