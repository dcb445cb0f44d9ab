"""Seeded fixtures for `junctions annotate` — TEST INFRASTRUCTURE.

`make_annotation_case(dir, seed)` writes a GTF, a BED12 of junctions and a FASTA whose shapes exercise every branch of
junctions_annotator.cc:128-311 and gtf_parser.cc: genes on both strands, 1-12 exons, single-exon genes (-S), several
transcripts per gene that share or shift exon ends (known donor / acceptor only, NDA), nested and overlapping genes,
introns from 70 b to 400 kb (bin levels 0-3 of the UCSC scheme), exon lines shuffled in the file, transcript ids in random
order, junctions that are known, exon-skipping, half-known, novel, on the other strand, with strand '?', at the last
exon's end (the reference reads past its exon vector there), on a contig the GTF lacks, and past the end of a FASTA
sequence (clipped fetch).  Everything is derived from random.Random(seed) and an explicit LCG, so it is reproducible.
"""
import os
import random

CONTIGS = [("1", 900000), ("10", 600000), ("2", 300000), ("3", 50000)]      # "3" has no genes


def write_fasta(path, contigs=CONTIGS, seed=7, width=70):
    x = (seed * 2654435761) & 0xFFFFFFFF
    alphabet = b"ACGTACGTACGTACGTACGTACGTACGTacgtN"
    with open(path, "wb") as f:
        for name, n in contigs:
            f.write(b">" + name.encode() + b" generated\n")
            buf = bytearray(n)
            for i in range(n):
                x = (x * 1664525 + 1013904223) & 0xFFFFFFFF
                buf[i] = alphabet[(x >> 24) % len(alphabet)]
            for i in range(0, n, width):
                f.write(bytes(buf[i:i + width]) + b"\n")
    return path


def _gene_models(rnd, contigs=None):
    """-> list of transcripts: dict(id, gene_name, gene_id, chrom, strand, exons=[(start, end)] ascending, 1-based inclusive)"""
    txs, tcount = [], 0
    for chrom, length in (contigs or CONTIGS)[:-1]:                             # the last contig carries no genes
        pos, g = 2000, 0
        while pos < length - 60000:
            g += 1
            strand = rnd.choice("+-")
            n_ex = rnd.choice([1, 1, 2, 3, 4, 5, 6, 8, 12])
            exons, p = [], pos
            for _ in range(n_ex):
                el = rnd.randrange(50, 300)
                exons.append((p, p + el - 1))
                il = rnd.choice([70, 90, 150, 400, 1200, 5000, 17000, 40000]) if rnd.random() < 0.93 else rnd.randrange(130000, 400000)
                p += el + il
            if exons[-1][1] >= length - 1000:
                break
            gene_id, gene_name = "G%s_%03d" % (chrom, g), rnd.choice(["GENE%s_%d" % (chrom, g), "SHARED", "A B"])
            variants = [exons]
            for _ in range(rnd.choice([0, 0, 1, 2, 3])):
                v = list(exons)
                if len(v) > 2 and rnd.random() < 0.6:
                    del v[rnd.randrange(1, len(v) - 1)]                      # skipped exon
                if len(v) > 1 and rnd.random() < 0.5:
                    k = rnd.randrange(len(v))
                    s, e = v[k]
                    v[k] = (s + rnd.choice([0, 3, 12]), e - rnd.choice([0, 4, 9]))   # alternative acceptor / donor
                if rnd.random() < 0.3:
                    v = v[:max(1, len(v) - 1)]
                variants.append(v)
            for v in variants:
                tcount += 1
                txs.append(dict(id="T%05d" % rnd.randrange(10 ** 5) + "_%d" % tcount, gene_name=gene_name, gene_id=gene_id, chrom=chrom,
                                strand=strand, exons=v))
            # next gene: sometimes nested / overlapping, usually downstream
            pos = rnd.choice([exons[0][0] + 20, exons[-1][1] + rnd.randrange(200, 30000), exons[-1][1] + rnd.randrange(200, 30000)])
    return txs


def write_gtf(path, txs, rnd):
    lines = []
    for t in txs:
        for k, (s, e) in enumerate(t["exons"]):
            attr = ('gene_id "%s"; transcript_id "%s"; exon_number "%d"; gene_name "%s"; gene_biotype "protein_coding";'
                    % (t["gene_id"], t["id"], k + 1, t["gene_name"]))
            lines.append("\t".join([t["chrom"], "gen", "exon", str(s), str(e), ".", t["strand"], ".", attr]))
            if rnd.random() < 0.3:
                lines.append("\t".join([t["chrom"], "gen", "CDS", str(s), str(e), ".", t["strand"], "0", attr]))
    rnd.shuffle(lines)
    with open(path, "w") as f:
        f.write("#!genome-build synthetic\n")
        f.write("\n".join(lines) + "\n")
    return path


def write_bed(path, txs, rnd, header=False, crlf=False, contigs=None):
    contigs = contigs or CONTIGS
    rows, n = [], 0

    def add(chrom, e_end, next_start, strand):
        nonlocal n
        if next_start - 1 <= e_end:
            return
        n += 1
        b0, b1 = rnd.randrange(8, 60), rnd.randrange(8, 60)
        s, e = e_end - b0, next_start - 1 + b1
        if s < 0:
            return
        rows.append([chrom, s, e, "JUNC%08d" % n, rnd.randrange(1, 500), strand, s, e, "255,0,0", 2, "%d,%d" % (b0, b1), "0,%d" % (e - s - b1)])

    for t in txs:
        ex, st, c = t["exons"], t["strand"], t["chrom"]
        for i in range(len(ex) - 1):
            r = rnd.random()
            if r < 0.5:
                add(c, ex[i][1], ex[i + 1][0], st)                                   # known junction
            if i + 2 < len(ex) and rnd.random() < 0.35:
                add(c, ex[i][1], ex[i + 2][0], st)                                   # skips an exon
            if rnd.random() < 0.2:
                add(c, ex[i][1], ex[i + 1][0] + rnd.choice([-7, 5, 33]), st)         # known donor side only
            if rnd.random() < 0.2:
                add(c, ex[i][1] + rnd.choice([-6, 4, 21]), ex[i + 1][0], st)         # known acceptor side only
            if rnd.random() < 0.1:
                add(c, ex[i][1], ex[i + 1][0], "-" if st == "+" else "+")            # other strand
            if rnd.random() < 0.05:
                add(c, ex[i][1], ex[i + 1][0], "?")
            if rnd.random() < 0.1:
                add(c, ex[i][1] + 11, ex[-1][0] + 17, st)                            # novel, spans several exons
        if rnd.random() < 0.3:
            add(c, ex[-1][1], ex[-1][1] + rnd.randrange(200, 90000), st)             # starts at the LAST exon's end
        if rnd.random() < 0.3:
            add(c, max(1, ex[0][0] - rnd.randrange(200, 50000)), ex[0][0], st)       # ends at the first exon's start
    for _ in range(40):
        c, ln = rnd.choice(contigs)
        a = rnd.randrange(100, ln - 2000)
        add(c, a, a + rnd.randrange(80, 1500), rnd.choice("+-"))                     # intergenic / contig without genes
    if contigs is CONTIGS:
        add("2", 299990, 300400, "+")                                                # past the end of the FASTA sequence (clipped)
    rnd.shuffle(rows)
    eol = "\r\n" if crlf else "\n"
    with open(path, "w", newline="") as f:
        if header:
            f.write("track name=junctions description=\"generated\"" + eol + "#comment" + eol)
        for r in rows:
            f.write("\t".join(str(x) for x in r) + eol)
    return path


def make_annotation_case(d, seed, header=False, crlf=False, contigs=None, fasta=True):
    """contigs: [(name, length)], the last one without genes; default = the small CONTIGS the goldens were made with."""
    os.makedirs(d, exist_ok=True)
    rnd = random.Random(seed)
    txs = _gene_models(rnd, contigs)
    gtf = write_gtf(os.path.join(d, "ann.gtf"), txs, rnd)
    bed = write_bed(os.path.join(d, "junctions.bed"), txs, rnd, header=header, crlf=crlf, contigs=contigs)
    fa = os.path.join(d, "ref.fa")
    if fasta and not os.path.exists(fa):
        write_fasta(fa, contigs or CONTIGS)
    return bed, fa, gtf


def write_fasta_numpy(path, contigs, seed=7, width=70):
    """Large genomes (bench): numpy instead of the per-base LCG; not byte-compatible with write_fasta."""
    import numpy as np
    rng = np.random.default_rng(seed)
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTACGTACGTACGTacgtN", dtype=np.uint8)
    with open(path, "wb") as f:
        for name, n in contigs:
            f.write(b">" + name.encode() + b" generated\n")
            full = (n // width) * width
            s = alphabet[rng.integers(0, len(alphabet), n)]
            body = np.empty((full // width, width + 1), dtype=np.uint8)
            body[:, :width] = s[:full].reshape(-1, width)
            body[:, width] = 10
            f.write(body.tobytes())
            if n > full:
                f.write(s[full:].tobytes() + b"\n")
    return path
