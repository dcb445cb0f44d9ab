"""Seeded synthetic SoA batches (numpy) for kernel-level parity tests."""
import numpy as np

OPS = "MIDNSHP=XB"


def cig(s):
    """'50M100N50M' -> uint32 words (len<<4|op)."""
    out, num = [], ""
    for ch in s:
        if ch.isdigit():
            num += ch
        else:
            out.append(int(num) << 4 | OPS.index(ch))
            num = ""
    return out


def batch_from_reads(reads):
    """reads: list of (tid, pos, flag, mapq, strand_byte, [cigar words])."""
    n = len(reads)
    tid = np.array([r[0] for r in reads], np.int32)
    pos = np.array([r[1] for r in reads], np.int32)
    meta = np.array([(r[2] << 16) | (r[3] << 8) | r[4] for r in reads], np.uint32)
    off = np.zeros(n + 1, np.uint32)
    words = []
    for i, r in enumerate(reads):
        off[i] = len(words)
        words.extend(r[5])
    off[n] = len(words)
    return tid, pos, meta, off, np.array(words, np.uint32) if words else np.zeros(0, np.uint32)


def random_batch(seed, n_reads, n_contigs=3, spliced_frac=0.15, weird=True, catalog_per_contig=40,
                 contig_len=2_000_000, read_len=100):
    """Coordinate-sorted reads drawn around a junction catalog so that keys merge heavily."""
    rng = np.random.default_rng(seed)
    cat = []
    for t in range(n_contigs):
        donors = np.sort(rng.integers(1000, contig_len - 700_000, catalog_per_contig))
        for d in donors:
            r = rng.random()
            if r < 0.05:
                ln = int(rng.integers(1, 70))            # below min intron
            elif r < 0.10:
                ln = int(rng.integers(500_001, 600_000))  # above max intron
            elif r < 0.14:
                ln = [69, 70, 500_000, 500_001][int(rng.integers(0, 4))]
            else:
                ln = int(np.exp(rng.uniform(np.log(70), np.log(100_000))))
            cat.append((t, int(d), ln, "+-"[int(rng.integers(0, 2))]))
    w = 1.0 / np.arange(1, len(cat) + 1)
    w = w[rng.permutation(len(cat))]
    w /= w.sum()
    flags = [0, 16, 99, 147, 83, 163, 65, 81, 97, 113, 129, 145, 161, 177, 4, 256 | 99, 1024 | 83, 2048, 512]
    reads = []
    for _ in range(n_reads):
        flag = flags[int(rng.integers(0, len(flags)))]
        mapq = [60, 255, 0, 1, 3][int(rng.integers(0, 5))]
        if rng.random() < spliced_frac:
            t, d, ln, st = cat[int(rng.choice(len(cat), p=w))]
            a = int(rng.integers(1, read_len))
            ops = []
            r = rng.random()
            left = a
            if weird and r < 0.10:
                s = int(rng.integers(1, 10)); ops.append(s << 4 | 4)            # leading S
            if weird and 0.10 <= r < 0.14:
                ops.append(5 << 4 | 5)                                            # leading H
            if weird and 0.14 <= r < 0.22 and left > 4:
                x = int(rng.integers(1, left - 1))
                mid = [2, 8, 1, 7, 6, 9, 11][int(rng.integers(0, 7))]           # D X I = P B op11
                ops += [x << 4 | 0, int(rng.integers(1, 4)) << 4 | mid, (left - x) << 4 | 0]
            else:
                ops.append(left << 4 | 0)
            ops.append(ln << 4 | 3)
            right = read_len - a
            r2 = rng.random()
            if r2 < 0.10 and right > 30:
                e = int(rng.integers(1, right - 1))
                ln2 = [100, 200, 50, 70][int(rng.integers(0, 4))]
                ops += [e << 4 | 0, ln2 << 4 | 3, (right - e) << 4 | 0]
            elif weird and r2 < 0.13:
                ops.append(int(rng.integers(70, 300)) << 4 | 3)                  # adjacent N N
                ops.append(right << 4 | 0)
            elif weird and r2 < 0.15:
                pass                                                              # trailing N
            elif weird and r2 < 0.20 and right > 6:
                e = int(rng.integers(1, right - 1))
                mid = [2, 8, 1, 4][int(rng.integers(0, 4))]
                ops += [e << 4 | 0, 2 << 4 | mid, (right - e) << 4 | 0]
            else:
                ops.append(right << 4 | 0)
            p = d - a
            rs = rng.random()
            sb = ord(st) if rs < 0.85 else (0 if rs < 0.93 else [ord("."), ord("*"), ord("?"), ord("x")][int(rng.integers(0, 4))])
            reads.append((t, p, flag, mapq, sb, ops))
        else:
            t = int(rng.integers(0, n_contigs))
            p = int(rng.integers(0, contig_len))
            r = rng.random()
            if r < 0.05:
                ops = [5 << 4 | 4, (read_len - 5) << 4 | 0]
            elif r < 0.08:
                ops = [40 << 4 | 0, 2 << 4 | 2, (read_len - 40) << 4 | 0]
            elif r < 0.10:
                ops = [40 << 4 | 0, 1 << 4 | 1, (read_len - 41) << 4 | 0]
            elif r < 0.11:
                ops = []                                                          # no CIGAR at all
            else:
                ops = [read_len << 4 | 0]
            reads.append((t, p, flag, mapq, 0, ops))
    if weird and n_reads > 10:
        reads[3] = (-1, reads[3][1], 4, 0, 0, cig("50M100N50M"))                  # tid -1 with a junction CIGAR
    reads.sort(key=lambda r: (r[0] if r[0] >= 0 else 1 << 30, r[1]))
    return batch_from_reads(reads)


def count_n_ops(cigar):
    return int(np.count_nonzero((cigar & 0xF) == 3))
