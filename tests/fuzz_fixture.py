"""Fuzzed BAM fixtures shared by the CPU differential tests (oracle vs the unmodified reference, tests/test_oracle.py) and the
GPU parity tests (product vs oracle, tests/test_gpu_zzzzz_fuzz.py) — TEST INFRASTRUCTURE."""
import os
import random
import struct
import subprocess

import bamio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CONTIGS = [("1", 20000000), ("10", 20000000), ("2", 20000000)]


def _index(bam):
    """The reference's own indexer when it was built (dev container; the binary travels to the GPU box), else bamgen's."""
    if os.path.exists(bam + ".bai"):
        os.remove(bam + ".bai")
    ref_index = os.path.join(ROOT, "oracle", "_ref", "ref_index")
    if os.path.exists(ref_index):
        subprocess.check_call([ref_index, bam])
    else:
        subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "index", bam], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return bam


def make_cigar_fuzz_bam(path, seed):
    """Every op incl. P / = / X / H, zero-length ops, N lengths on both QC bounds, up to 20 ops, N first or last, flags incl.
    unmapped-with-CIGAR, XS of type A / C / absent / twice."""
    rnd = random.Random(9000 + seed)
    reads = []
    for _ in range(rnd.choice([3, 30, 300])):
        cig = []
        for _k in range(rnd.choice([1, 2, 3, 3, 4, 5, 8, 20])):
            op = rnd.choice("MMMMNNNIDSH=XP")
            ln = rnd.choice([0, 1, 5, 50, 69, 70, 100, 500000, 500001]) if op == "N" else rnd.choice([0, 1, 3, 7, 8, 20, 50])
            cig.append((ln << 4) | bamio.OPS.index(op))
        reads.append((rnd.choice([0, 0, 1, 2]), rnd.randrange(0, 3000), cig, rnd.choice([0, 16, 4, 99, 147, 83, 163, 256, 1024, 2048 + 16]),
                      rnd.choice([0, 1, 60, 255]), rnd.choice([b"XSA+", b"XSA-", b"XSA.", b"", b"XSC\x2b", b"NHC\x01XSA-", b"XSA+XSA-"])))
    reads.sort(key=lambda x: (x[0], x[1]))
    recs = [bamio.record(t, p, c, f, q, a, name=b"q%05d" % i, l_seq=10) for i, (t, p, c, f, q, a) in enumerate(reads)]
    bamio.write_bam(path, CONTIGS, recs, block_size=rnd.choice([0x300, 0xff00]))
    return _index(path)


CIGAR_FUZZ_ARGS = (["-s", "XS"], ["-s", "RF", "-a", "0", "-m", "0", "-M", "4000000000"], ["-s", "FR", "-a", "1", "-m", "1"],
                   ["-s", "XS", "-r", "1:100-2000"], ["-s", "XS", "-t", "NH"])


def make_barcode_fuzz_bam(path, seed):
    """CB as Z or H, before / after / between XS, NH, Z, B and f tags, barcodes of 1-40 characters incl. ':' ',' '-', 1-200
    distinct barcodes, reads without CB, XS of type Z (-> '?'), small BGZF blocks."""
    rnd = random.Random(5000 + seed)
    bcs = [("".join(rnd.choice("ACGT:,-_") for _ in range(rnd.choice([1, 4, 16, 40])))).encode() for _ in range(rnd.choice([1, 3, 20, 200]))]
    loci = [(rnd.choice([0, 1, 2]), rnd.randrange(1000, 50000, 1000),
             rnd.choice(["50M100N50M", "20M300N30M500N50M", "5S45M1000N50M", "50M69N50M", "3M200N97M", "50M2D50M", "100M"]))
            for _ in range(rnd.randrange(1, 8))]
    reads = []
    for _ in range(rnd.choice([5, 50, 400])):
        tid, pos, cg = rnd.choice(loci)
        xs = rnd.choice([b"XSA+", b"XSA-", b"XSA?", b"", b"XSZ+\0"])
        r = rnd.random()
        cb = b"" if r < 0.1 else (b"CBH" if r < 0.2 else b"CBZ") + rnd.choice(bcs) + b"\0"
        other = rnd.choice([b"", b"NHC\x01", b"RGZx y\0", b"ZBBc" + struct.pack("<I", 2) + b"\x01\x02", b"XXf" + struct.pack("<f", 1.5)])
        parts = [xs, cb, other]
        rnd.shuffle(parts)
        reads.append((tid, pos, cg, rnd.choice([0, 16, 99, 147]), b"".join(parts)))
    reads.sort(key=lambda x: (x[0], x[1]))
    recs = [bamio.record(t, p, c, f, 60, a, name=b"q%05d" % i) for i, (t, p, c, f, a) in enumerate(reads)]
    bamio.write_bam(path, [("1", 100000), ("10", 100000), ("2", 100000)], recs, block_size=rnd.choice([0x200, 0x4000, 0xff00]))
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "index", path], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return path


BARCODE_FUZZ_ARGS = (["-s", "XS"], ["-s", "RF", "-a", "3"], ["-s", "XS", "-r", "10"])


def make_motif_fuzz_case(d, seed):
    """Intron-motif strand mode: a genome made of the six motif dimers (so GT-AG / CT-AC ... hits are common) with some lower
    case and N, contig 2 shorter than its junctions (clipped fetch), contig 10 sometimes absent from the FASTA (the reference
    throws), alignments with 1-3 junctions (the reused-Junction reverse-complement quirk).  -> (bam, fasta)"""
    rnd = random.Random(4000 + seed)

    def genome(n):
        return "".join(rnd.choice(["GT", "AG", "GC", "AT", "AC", "CT", "gt", "NN"] if rnd.random() < 0.1 else ["GT", "AG", "GC", "AT", "AC", "CT"])
                       for _ in range(n // 2))
    contigs = [("1", 6000), ("10", 6000), ("2", 6000)]
    fa = os.path.join(d, "g.fa")
    drop10 = rnd.random() < 0.1
    with open(fa, "w") as f:
        for name, n in contigs:
            if name == "10" and drop10:
                continue
            s = genome(n if name != "2" else 1500)
            f.write(f">{name} d\n" + "\n".join(s[i:i + 50] for i in range(0, len(s), 50)) + "\n")
    if os.path.exists(fa + ".fai"):
        os.remove(fa + ".fai")
    reads = []
    for _ in range(rnd.choice([5, 60, 300])):
        cig = [(rnd.choice([10, 11, 20, 21]) << 4) | 0]
        for _k in range(rnd.choice([1, 1, 2, 3])):
            cig.append((rnd.choice([70, 71, 100, 101, 200]) << 4) | 3)
            cig.append((rnd.choice([10, 11, 30, 31]) << 4) | 0)
        reads.append((rnd.choice([0, 1, 2]), rnd.randrange(0, 3000), cig, rnd.choice([0, 16, 99, 147]), 60, rnd.choice([b"XSA+", b"XSA-", b"", b"XSA?"])))
    reads.sort(key=lambda x: (x[0], x[1]))
    recs = [bamio.record(t, p, c, f, q, a, name=b"q%05d" % i, l_seq=10) for i, (t, p, c, f, q, a) in enumerate(reads)]
    bam = os.path.join(d, "f.bam")
    bamio.write_bam(bam, contigs, recs)
    return _index(bam), fa


MOTIF_FUZZ_ARGS = (["-s", "XS"], ["-s", "RF"], ["-s", "intron-motif"], ["-s", "FR", "-a", "0"])


def damage_bam(src, dst, seed):
    """A copy of `src` (+ its .bai) with one kind of damage: truncation, 1-3 flipped bits, 8 zeroed bytes, or a wrong ISIZE
    trailer.  -> the kind"""
    import shutil
    rnd = random.Random(seed)
    data = bytearray(open(src, "rb").read())
    mode = rnd.choice(["trunc", "flip", "zero", "isize"])
    if mode == "trunc":
        data = data[:rnd.randrange(len(data) // 3, len(data))]
    elif mode == "flip":
        for _ in range(rnd.choice([1, 3])):
            data[rnd.randrange(200, len(data))] ^= 1 << rnd.randrange(8)
    elif mode == "zero":
        i = rnd.randrange(200, len(data) - 8)
        data[i:i + 8] = bytes(8)
    else:
        off, blocks = 0, []
        while off + 18 <= len(data):
            bs = struct.unpack_from("<H", data, off + 16)[0] + 1
            blocks.append((off, bs))
            off += bs
        o, bs = rnd.choice(blocks[:-1])
        data[o + bs - 4:o + bs] = struct.pack("<I", rnd.choice([0, 1, 70000, struct.unpack_from("<I", data, o + bs - 4)[0] + 1]))
    with open(dst, "wb") as f:
        f.write(data)
    shutil.copy(src + ".bai", dst + ".bai")
    return mode
