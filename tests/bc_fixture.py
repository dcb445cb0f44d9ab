"""Seeded single-cell (`-b`) BAM fixtures — TEST INFRASTRUCTURE.

`make_barcode_bam` writes a small coordinate-sorted, indexed BAM whose alignments carry CB:Z barcodes in the shapes that
matter to set_junction_barcode / add_junction / print_barcodes (junctions_extractor.cc:203-215,362-374, .h:99-111):
several contigs (string order != tid order), a hot junction shared by hundreds of barcodes (several rehashes of the
reference's unordered_map), junctions seen by one barcode, the same barcode on many junctions, alignments with two
junctions, alignments without CB ("?" + a WARNING line), CB before / after XS and behind Z and B tags, proxy-2 strands
('?', '.', no XS), junctions that fail junction_qc (never stored) and junctions hidden by the anchor filter (stored, no
barcode line printed).
"""
import os
import random
import struct
import subprocess

import bamio

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BAMGEN = os.path.join(ROOT, "tools", "bamgen")

CONTIGS = [("1", 400000), ("10", 400000), ("2", 400000)]


def make_barcode_bam(path, seed=11, n_reads=3000, n_barcodes=400, hot_barcodes=300, missing=0.04):
    rnd = random.Random(seed)
    bcs = [("".join(rnd.choice("ACGT") for _ in range(16)) + "-1").encode() for _ in range(n_barcodes)]
    # loci: (tid, pos, cigar, weight, barcode pool size)
    loci = []
    for tid in range(len(CONTIGS)):
        loci.append((tid, 10000, "50M2000N50M", 30, hot_barcodes))          # hot junction: many barcodes
        loci.append((tid, 10010, "40M2000N60M", 10, hot_barcodes))          # same junction key, other anchors
        loci.append((tid, 30000, "30M500N20M700N50M", 8, 40))               # two junctions per alignment
        loci.append((tid, 50000, "5M300N95M", 3, 10))                       # hidden by -a 8 unless another read anchors it
        loci.append((tid, 49990, "15M300N85M", 1, 10))                      # ... this one does
        loci.append((tid, 70000, "50M69N50M", 2, 10))                       # fails junction_qc (69 < 70)
        loci.append((tid, 90000, "5S45M1200N50M", 4, 3))
        loci.append((tid, 110000, "50M", 6, 50))                            # n_cigar == 1: set_junction_barcode never runs
        loci.append((tid, 130000, "50M2I50M", 3, 50))                       # n_cigar > 1, no junction: warning if CB is missing
        for k in range(12):
            loci.append((tid, 150000 + 3000 * k, "%dM%dN%dM" % (20 + k, 100 + 37 * k, 80 - k), 2, 1 + 5 * k))
    weights = [l[3] for l in loci]
    reads = []
    for i in range(n_reads):
        tid, pos, cg, _, pool = rnd.choices(loci, weights)[0]
        pos += rnd.randrange(0, 3) * 0                                      # fixed positions keep the keys shared
        r = rnd.random()
        strand_aux = rnd.choice([b"XSA+", b"XSA+", b"XSA-", b"XSA?", b"XSA.", b""])
        if r < missing:
            aux = strand_aux + b"NHC\x01"
        else:
            bc = rnd.choice(bcs[:pool])
            shape = rnd.randrange(5)
            cb = b"CBZ" + bc + b"\0"
            if shape == 0:
                aux = cb + strand_aux
            elif shape == 1:
                aux = strand_aux + cb
            elif shape == 2:
                aux = b"RGZgroup1\0" + strand_aux + b"NHC\x02" + cb
            elif shape == 3:
                aux = b"ZBBS" + struct.pack("<I", 3) + struct.pack("<3H", 1, 2, 3) + cb + strand_aux
            else:
                aux = b"ASi" + struct.pack("<i", 77) + cb + b"CBZSECOND-IGNORED\0" + strand_aux    # first match wins
        flag = rnd.choice([0, 16, 99, 147])
        reads.append((tid, pos, cg, flag, aux))
    reads.sort(key=lambda x: (x[0], x[1]))
    recs = [bamio.record(t, p, c, f, 60, a, name=b"q%05d" % i) for i, (t, p, c, f, a) in enumerate(reads)]
    bamio.write_bam(path, CONTIGS, recs, block_size=0x4000)
    subprocess.check_call([BAMGEN, "index", path])
    return path
