"""CPU: host-side logic of the product — native BGZF/BAM/BAI feeder, option parsing, shard planning,
table import/merge/naming/BED12 formatting — checked against the oracle and the reference's goldens."""
import io
import os
import subprocess

import numpy as np
import pytest

import bamio
from oracle_py import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
HCC = os.path.join(GOLD, "hcc1395", "test_hcc1395.bam")
KAT = os.path.join(GOLD, "kat", "kat.bam")
SYN = os.path.join(GOLD, "kat", "synth.bam")
BAMGEN = os.path.join(ROOT, "tools", "bamgen")


def rt():
    import regtools_b200
    return regtools_b200


def feeder_vs_oracle(bam, region, strandness=0, tag="XS", threads=3):
    """The product's feeder output (SoA batch), pushed through the oracle's per-read walk, must reproduce
    the oracle's own BAM reader exactly: same reads, same order, same strand bytes."""
    ex = rt().JunctionsExtractor(bam, region, strandness, tag, 8, 70, 500000, device=-1, n_threads=threads)
    arrs = ex.load_batch()
    names = ex.contig_names()
    ex.close()
    a = Oracle(8, 70, 500000, strandness, tag, contigs=names)
    a.batch(*arrs)
    b = Oracle(8, 70, 500000, strandness, tag)
    b.extract_bam(bam, region)
    assert len(arrs[0]) == b.reads_seen()
    assert a.bed12() == b.bed12()
    ta, tb = a.table(), b.table()
    assert np.array_equal(ta, tb)
    return arrs


@pytest.mark.parametrize("region", [".", "1:22405013-22405020", "1", "1:22,400,000-22,410,000", "1:22405013", "*"])
def test_feeder_hcc(region):
    if region == "*":
        with pytest.raises(RuntimeError):      # no unplaced reads indexed -> iterator NULL in the reference
            feeder_vs_oracle(HCC, region)
        return
    feeder_vs_oracle(HCC, region)


@pytest.mark.parametrize("bam", [KAT, SYN])
@pytest.mark.parametrize("region,strandness,tag", [(".", 0, "XS"), (".", 1, "XS"), ("10", 0, "XS"), ("2:100-5000100", 0, "XS"),
                                                   (".", 0, "NH"), ("1:5000-6200", 2, "XS")])
def test_feeder_kat_and_synth(bam, region, strandness, tag):
    feeder_vs_oracle(bam, region, strandness, tag)


def test_feeder_thread_counts_agree():
    ref = None
    for n in (1, 2, 7):
        arrs = feeder_vs_oracle(SYN, ".", threads=n)
        if ref is None:
            ref = arrs
        else:
            assert all(np.array_equal(x, y) for x, y in zip(arrs, ref))


def test_feeder_large_generated_bam(tmp_path):
    bam = str(tmp_path / "g.bam")
    subprocess.check_call([BAMGEN, "gen", "--out", bam, "--config", "tiny", "--reads", "150000", "--seed", "5", "--threads", "4"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    arrs = feeder_vs_oracle(bam, ".", threads=4)
    assert len(arrs[0]) == 150000
    feeder_vs_oracle(bam, "10:100000-1500000")


def test_empty_block_in_the_middle_ends_iteration(tmp_path):
    """htslib 1.2.1 bgzf_read breaks on an ISIZE=0 block (bgzf.c:559-561): records after it are never seen."""
    recs1 = [bamio.record(0, 100 + i, "50M100N50M", aux=b"XSA+") for i in range(5)]
    recs2 = [bamio.record(0, 5000 + i, "50M200N50M", aux=b"XSA-") for i in range(5)]
    good = str(tmp_path / "good.bam")
    bamio.write_bam(good, [("c", 100000)], recs1 + recs2)
    subprocess.check_call([BAMGEN, "index", good], stderr=subprocess.DEVNULL)
    bad = str(tmp_path / "bad.bam")
    # header block + records1 + EMPTY + records2 + EOF, reusing the index of the good file's layout is
    # not possible (offsets move), so index the bad file with the first part only and append afterwards
    bamio.write_bam(bad, [("c", 100000)], recs1)
    subprocess.check_call([BAMGEN, "index", bad], stderr=subprocess.DEVNULL)
    with open(bad, "r+b") as f:
        data = f.read()
        assert data.endswith(bamio.EOF_BLOCK)
        f.seek(len(data))            # keep the first EOF block as the "empty block in the middle"
        f.write(bamio.bgzf_block(b"".join(recs2)) + bamio.EOF_BLOCK)
    arrs = feeder_vs_oracle(bad, ".")
    assert len(arrs[0]) == 5


def test_truncated_file_ends_silently(tmp_path):
    recs = [bamio.record(0, 100 + 10 * i, "50M100N50M", aux=b"XSA+") for i in range(2000)]
    full = str(tmp_path / "full.bam")
    bamio.write_bam(full, [("c", 1000000)], recs, block_size=4000)
    subprocess.check_call([BAMGEN, "index", full], stderr=subprocess.DEVNULL)
    size = os.path.getsize(full)
    with open(full, "r+b") as f:
        f.truncate(size * 2 // 3)
    arrs = feeder_vs_oracle(full, ".")
    assert 0 < len(arrs[0]) < 2000


def test_error_messages_and_exit_codes(tmp_path, capsys):
    """tests/integration-test/test_junctions_extract.py:87-109 + the runtime_error texts of junctions_extractor.cc."""
    r = rt()
    out = str(tmp_path / "o.bed")
    assert r.junctions_extract(["extract", "-s", "XS", "-o", out]) == 1
    assert r.junctions_extract(["extract", "-s", "XS", "-o", out, "does_not_exist.bam"]) == 1
    assert "Unable to open BAM/SAM file." in capsys.readouterr().err
    assert r.junctions_extract(["extract", "-o", out, HCC]) == 1
    assert "Please supply strandness mode" in capsys.readouterr().err
    assert r.junctions_extract(["extract", "-s", "bogus", HCC]) == 1
    assert r.junctions_extract(["extract", "-h"]) == 0
    assert r.junctions_extract(["extract", "-s", "intron-motif", HCC]) == 1
    noidx = str(tmp_path / "noidx.bam")
    bamio.write_bam(noidx, [("c", 1000)], [bamio.record(0, 1, "10M")])
    assert r.junctions_extract(["extract", "-s", "XS", noidx]) == 1
    assert "Unable to open BAM/SAM index" in capsys.readouterr().err
    ex = r.JunctionsExtractor(HCC, "nonexistent_contig:1-10", 0, device=-1)
    with pytest.raises(RuntimeError, match="Unable to iterate to region"):
        ex.load_batch()
    ex = r.JunctionsExtractor(HCC, "1:100-50", 0, device=-1)
    with pytest.raises(RuntimeError, match="Unable to iterate to region"):
        ex.load_batch()


def test_parse_options_defaults_and_atoi():
    r = rt()
    ex = r.JunctionsExtractor()
    ex.parse_options(["extract", "-s", "RF", "-a", "12junk", "-m", "x", "-M", "-1", "-r", "1:5-9", "-t", "ZS", "-o", "f", "a.bam"])
    assert (ex.min_anchor_length_, ex.min_intron_length_, ex.max_intron_length_) == (12, 0, 0xFFFFFFFF)
    assert (ex.strandness_, ex.region_, ex.strand_tag_, ex.output_file_, ex.get_bam()) == (1, "1:5-9", "ZS", "f", "a.bam")
    ex = r.JunctionsExtractor.from_region("a.bam", "1:1-2", 0, "XS", 8, 70, 500000)
    assert ex.min_intron_length_ == 8        # junctions_extractor.h:199-200


def test_plan_shards_balances_contigs():
    assign = rt().plan_shards(SYN, 2)
    assert len(assign) == 3 and set(assign) == {0, 1}
    assert rt().plan_shards(SYN, 1) == [0, 0, 0]
    assert len(set(rt().plan_shards(SYN, 8))) == 3


def test_plan_shards_are_contiguous_runs_of_contigs(tmp_path):
    """A shard is a run of consecutive contigs (one contiguous byte range of the sorted file per rank), balanced on
    compressed bytes: min-max over contiguous partitions, spare ranks split the heaviest runs."""
    bam = str(tmp_path / "wg.bam")
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "gen", "--out", bam, "--config", "c3", "--reads", "120000", "--seed", "3",
                           "--threads", "2"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sizes = None
    for world in (2, 3, 4, 8, 24, 40):
        a = rt().plan_shards(bam, world)
        assert len(a) == 24 and a == sorted(a) and a[0] == 0                     # runs of consecutive contigs, in file order
        assert set(a) == set(range(min(world, 24)))                              # every rank that can get a contig gets one
        if sizes is None:
            ex = rt().JunctionsExtractor(bam, ".", 0, device=-1)
            tid = ex.load_batch()[0]
            ex.close()
            sizes = np.bincount(tid[tid >= 0], minlength=24).astype(float)
        load = np.array([sizes[[t for t in range(24) if a[t] == r]].sum() for r in range(min(world, 24))])
        # no contiguous partition can beat max(largest contig, average); allow 35 % on top (reads ~ bytes only roughly)
        assert load.max() <= 1.35 * max(sizes.max(), sizes.sum() / min(world, 24)), (world, load)


def _oracle_table_as_rtjx(o, first_ord_base=0):
    r = rt()
    t = o.table()
    out = np.zeros(len(t), r.JUNCTION_DTYPE)
    for f in ("tid", "start", "end", "thick_start", "thick_end", "read_count", "name_index", "strand", "left_ok", "right_ok"):
        out[f] = t[f]
    out["first_ord"] = t["name_index"].astype(np.uint64) + first_ord_base   # monotone in first appearance
    return out


def test_import_merge_renames_by_tid_then_first_seen():
    """Per-contig shard tables imported in arbitrary order give the whole-file BED12."""
    r = rt()
    whole = Oracle(8, 70, 500000, 0)
    whole.extract_bam(SYN)
    names = [whole.l.jxo_contig(whole.h, i).decode() for i in range(3)]
    shards = []
    for contig in ("2", "1", "10"):                       # deliberately not in tid order
        o = Oracle(8, 70, 500000, 0)
        o.extract_bam(SYN, contig)
        shards.append(_oracle_table_as_rtjx(o))
    merged = r.JunctionsExtractor(SYN, device=-1)
    merged.set_contigs(names)
    for s in shards:
        merged.import_table(s)
    buf = io.StringIO()
    merged.print_all_junctions(buf)
    assert buf.getvalue() == whole.bed12()
    got = merged.junction_table()
    want = whole.table()
    for f in ("tid", "start", "end", "read_count", "name_index", "strand"):
        assert np.array_equal(got[f], want[f])


def test_name_index_above_1e8_sorts_as_string():
    """JUNC%08d wider than 8 digits: compare_junctions falls back to std::string order (junctions_extractor.h:139)."""
    r = rt()
    t = np.zeros(3, r.JUNCTION_DTYPE)
    t["tid"] = 0; t["start"] = 100; t["end"] = 200; t["thick_start"] = 50; t["thick_end"] = 250
    t["read_count"] = 1; t["strand"] = ord("+"); t["left_ok"] = 1; t["right_ok"] = 1
    t["end"] = [200, 201, 202]
    h = r.JunctionsExtractor(device=-1)
    h.set_contigs(["c"])
    h.import_table(t)
    assert [int(x) for x in h.junction_table()["name_index"]] == [1, 2, 3]



def test_sorted_shard_tables_merge_without_sorting(capfd, monkeypatch):
    """Rank 0 of a multi-GPU run: shard tables (runs of consecutive contigs, sorted, names ranked inside the shard) are
    merged in O(n); tables that break an assumption take the general re-rank + sort path.  Both must print the whole
    file's BED12."""
    monkeypatch.setenv("RTJX_TRACE", "1")
    r = rt()
    whole = Oracle(8, 70, 500000, 0)
    whole.extract_bam(SYN)
    want = whole.bed12()

    def shard(region, tid_ord_base):
        o = Oracle(8, 70, 500000, 0)
        o.extract_bam(SYN, region)
        t = o.table()
        part = np.zeros(len(t), r.JUNCTION_DTYPE)
        for f in ("tid", "start", "end", "thick_start", "thick_end", "read_count", "name_index", "strand", "left_ok", "right_ok"):
            part[f] = t[f]
        part["first_ord"] = tid_ord_base + t["name_index"].astype(np.uint64)      # monotone in the shard's name order
        return part

    parts = [shard("1", 0), shard("10", 1 << 32), shard("2", 2 << 32)]
    for order, expect in (((0, 1, 2), "merged without sorting"), ((2, 0, 1), "merged without sorting")):
        ex = r.JunctionsExtractor(SYN, ".", 0, device=-1)
        ex.set_contigs(["1", "10", "2"])
        for k in order:
            ex.import_table(parts[k])
        buf = io.StringIO()
        ex.print_all_junctions(buf)
        ex.close()
        assert buf.getvalue() == want
        assert expect in capfd.readouterr().err
    broken = [p.copy() for p in parts]
    broken[1]["name_index"][:] = broken[1]["name_index"][::-1]                   # names no longer follow (tid, first_ord)
    ex = r.JunctionsExtractor(SYN, ".", 0, device=-1)
    ex.set_contigs(["1", "10", "2"])
    for p in broken:
        ex.import_table(p)
    buf = io.StringIO()
    ex.print_all_junctions(buf)
    ex.close()
    assert buf.getvalue() == want and "general path" in capfd.readouterr().err


CSI_CASES = [("kat.bam", "kat.min14.csi"), ("kat.bam", "kat.min12.csi"), ("synth.bam", "synth.min14.csi")]


def _with_csi(tmp_path, bam, csi, also_bai=False):
    """Copy of a fixture BAM that has only a .csi next to it (built by the reference's htslib: oracle/ref_index.c)."""
    import shutil
    dst = str(tmp_path / bam)
    shutil.copy(os.path.join(GOLD, "kat", bam), dst)
    shutil.copy(os.path.join(GOLD, "csi", csi), dst + ".csi")
    if also_bai:
        shutil.copy(os.path.join(GOLD, "kat", bam + ".bai"), dst + ".bai")
    return dst


@pytest.mark.parametrize("bam,csi", CSI_CASES)
def test_csi_index_gives_the_same_alignments_as_bai(bam, csi, tmp_path):
    """hts_idx_load prefers <bam>.csi (hts.c:2031-2042); a CSI is BGZF-compressed, has its own min_shift / depth, a loff
    per bin and no linear index (hts.c:1580-1594, 1535-1538).  Every region must yield exactly the BAI's alignments."""
    src = _with_csi(tmp_path, bam, csi)
    regions = [".", "1", "10", "2", "1:5000-6200", "1:900-1300", "10:500000-900000", "2:5,000,050-5,000,060", "1:1-1", "2:2400000-2500000"]
    for reg in regions:
        a = rt().JunctionsExtractor(src, reg, 0, device=-1)
        b = rt().JunctionsExtractor(os.path.join(GOLD, "kat", bam), reg, 0, device=-1)
        xa, xb = a.load_batch(), b.load_batch()
        a.close(); b.close()
        assert all(np.array_equal(u, v) for u, v in zip(xa, xb)), (csi, reg)
        if reg == ".":
            assert len(xa[0]) > 40
    assert rt().plan_shards(src, 2) == rt().plan_shards(os.path.join(GOLD, "kat", bam), 2)


def test_csi_takes_precedence_over_bai(tmp_path):
    src = _with_csi(tmp_path, "kat.bam", "kat.min12.csi", also_bai=True)
    open(src + ".bai", "wb").write(b"not an index")                 # would fail if it were read
    ex = rt().JunctionsExtractor(src, "1:5000-6200", 0, device=-1)
    assert len(ex.load_batch()[0]) > 0
    ex.close()


# ---- `-b` single-cell mode: the host feeder's barcode column ---------------------------------------------------------
def _py_barcodes(bam):
    """Independent walk of the BAM (zlib + struct): per alignment (n_cigar, first CB:Z value or None)."""
    import struct
    import zlib
    raw, data, o = open(bam, "rb").read(), b"", 0
    while o + 18 <= len(raw):
        bs = struct.unpack_from("<H", raw, o + 16)[0] + 1
        data += zlib.decompress(raw[o + 18:o + bs - 8], -15)
        o += bs
    p = 8 + struct.unpack_from("<i", data, 4)[0]
    n_ref = struct.unpack_from("<i", data, p)[0]; p += 4
    for _ in range(n_ref):
        p += 8 + struct.unpack_from("<i", data, p)[0]
    size = {b"A": 1, b"c": 1, b"C": 1, b"s": 2, b"S": 2, b"i": 4, b"I": 4, b"f": 4, b"d": 8}
    out = []
    while p + 4 <= len(data):
        bl = struct.unpack_from("<i", data, p)[0]
        _tid, _pos, l_rn, _mq, _bin, n_cig, _flag, l_seq = struct.unpack_from("<iiBBHHHi", data, p + 4)
        a = p + 36 + l_rn + 4 * n_cig + (l_seq + 1) // 2 + l_seq
        e, bc = p + 4 + bl, None
        while a + 3 <= e:
            tag, ty = data[a:a + 2], data[a + 2:a + 3]
            a += 3
            if ty in (b"Z", b"H"):
                z = data.index(b"\0", a)
                if tag == b"CB":
                    bc = data[a:z].decode(); break
                a = z + 1
            elif ty == b"B":
                sub, n = data[a:a + 1], struct.unpack_from("<I", data, a + 1)[0]
                a += 5 + size[sub] * n
            else:
                if tag == b"CB":
                    break
                a += size[ty]
        out.append((n_cig, bc))
        p += 4 + bl
    return out


@pytest.mark.parametrize("threads", [1, 4])
def test_feeder_barcode_column(threads, tmp_path):
    """rtjx_load_barcodes (host only): every n_cigar > 1 alignment carries the dictionary id of its first CB:Z value, "?"
    when the tag is absent (set_junction_barcode, junctions_extractor.cc:362-374); ids are first-seen ranks."""
    import bc_fixture
    import regtools_b200 as rt
    for bam in (os.path.join(ROOT, "tests", "golden", "barcodes", "bc2.bam"),
                bc_fixture.make_barcode_bam(str(tmp_path / "f.bam"), seed=5, n_reads=1500, missing=0.2)):
        ex = rt.JunctionsExtractor(bam, ".", 0, device=-1, n_threads=threads, batch_reads=1024)
        ex.output_barcodes_file_ = os.devnull
        ids = ex.load_barcodes()
        names = ex.barcode_names()
        n_bc, n_missing = ex.barcode_stats()
        ex.close()
        want = _py_barcodes(bam)
        assert len(ids) == len(want) and n_bc == len(names) == len(set(names))
        seen, missing = [], 0
        for i, (n_cig, bc) in enumerate(want):
            if n_cig <= 1:
                assert ids[i] == 0
                continue
            if bc is None:
                missing += 1
                bc = "?"
            if bc not in seen:
                seen.append(bc)
            assert names[ids[i]] == bc, i
        assert names == seen and n_missing == missing and missing > 0


# ---- FASTA loader (fasta.cc): several threads, same result as one sequential faidx-style pass -------------------------
def _py_fasta(path):
    """faidx's view (fai_build_core, faidx.c:82-155): a record starts at '>' at the beginning of a line, the name runs to the
    first white space, the bases are the isgraph() bytes of the following lines, a repeated name is ignored."""
    seqs, order, cur = {}, [], None
    for line in open(path, "rb").read().split(b"\n"):
        if line.startswith(b">"):
            name = line[1:].lstrip(b" \t\r\v\f").split(None, 1)
            name = name[0] if name else b""
            cur = None
            if name not in seqs:
                seqs[name] = bytearray(); order.append(name); cur = name
        elif cur is not None:
            seqs[cur] += bytes(c for c in line if 32 < c < 127)
    return [(n, bytes(seqs[n])) for n in order]


def _fnv(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_fasta_loader_matches_a_sequential_reading(tmp_path):
    import random
    import subprocess
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "-s", "../../build/emul/fasta_dump"])
    exe = os.path.join(ROOT, "build", "emul", "fasta_dump")
    rnd = random.Random(3)
    cases = {}
    cases["plain"] = b">1 desc\nACGT\nacgtN\n>2\nGG\n"
    cases["crlf_noeol"] = b">a\r\nAC GT\r\nNN\r\n>b x y\r\nTTTT"
    cases["dups_junk"] = b"junk before\nmore\n>x\nAAAA\n>y\nCC\n>x\nGGGG\n>z\n\n\nT>T\n>\nAA\n"
    big = bytearray()
    for i in range(40):                                        # ~14 MB: several chunks, headers and duplicates inside chunks
        big += b">seq%d some text\n" % (i % 33)
        for _ in range(rnd.randrange(1, 6000)):
            big += bytes(rnd.choice(b"ACGTacgtN") for _ in range(60)) + rnd.choice([b"\n", b"\n", b" \n", b"\r\n"])
    cases["big"] = bytes(big)
    for name, data in cases.items():
        path = tmp_path / (name + ".fa")
        path.write_bytes(data)
        got = subprocess.run([exe, str(path)], capture_output=True, text=True, check=True).stdout.splitlines()
        want = _py_fasta(str(path))
        assert len(got) == len(want), name
        for line, (n, s) in zip(got, want):
            f = line.split("\t")
            assert f[0] == n.decode() and int(f[1]) == len(s) and int(f[2], 16) == _fnv(s), (name, n)


@pytest.mark.parametrize("block", range(3))
def test_feeder_on_fuzzed_bams(block, tmp_path):
    """The product's host feeder on the fuzzed BAMs of tests/fuzz_fixture.py (random CIGARs incl. zero-length and P ops,
    unmapped-with-CIGAR flags, strand tags of several types, tiny BGZF blocks; aux layouts with Z / H / B / f tags): its SoA
    batch pushed through the oracle's walk must equal the oracle's own reader — which the CPU suite pins to the unmodified
    reference on the same generators.  Whole file, a region, and NH as the strand tag; 1 and 3 inflate threads."""
    import fuzz_fixture as ff
    for seed in range(block * 8, block * 8 + 8):
        bam = ff.make_cigar_fuzz_bam(str(tmp_path / "c.bam"), seed)
        for region, strandness, tag, threads in ((".", 0, "XS", 3), ("1:100-2000", 0, "XS", 1), (".", 1, "XS", 1), (".", 0, "NH", 3), ("2", 2, "XS", 3)):
            feeder_vs_oracle(bam, region, strandness, tag, threads)
        bam = ff.make_barcode_fuzz_bam(str(tmp_path / "b.bam"), seed)
        for region in (".", "10"):
            feeder_vs_oracle(bam, region, 0, "XS", 2)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "regtools_ref")), reason="needs oracle/_ref/regtools_ref (dev container)")
def test_odd_region_strings_three_way():
    """hts_parse_reg / hts_parse_decimal corner cases (hts.c:1833-1922): exit code and BED12 of the unmodified reference, the
    oracle's reader and the product's feeder (+ oracle walk) must agree on every spelling."""
    ref = os.path.join(ROOT, "oracle", "_ref", "regtools_ref")
    regions = ["1", "1:", "1:5000", "1:5000-", "1:-6200", "1:5,000-6,200", "1:5000-6200", "1:6200-5000", "1:0-100", "1:1e3-9e3", "1:1.5e3-9e3",
               "1:5000-6200 ", "10", "10:1", "2:1-1", "2:0", "x", "1:abc", "1:5000-abc", ":", "", "1:5000-6200:7", "1:-", "1:5000--6200", "01",
               "1:+5000-6200", "1:5000-6200k", "1:5k-7k", "1:0.005M-0.007M", "1:1-3G", "*", "."]
    for reg in regions:
        p = subprocess.run([ref, "junctions", "extract", "-s", "XS", "-a", "0", "-m", "0", "-r", reg, KAT], capture_output=True, text=True)
        o = Oracle(0, 0, 500000, 0)
        try:
            o.extract_bam(KAT, reg)
            orc, ob = 0, o.bed12()
        except RuntimeError:
            orc, ob = 1, None
        try:
            ex = rt().JunctionsExtractor(KAT, reg, 0, "XS", 0, 0, 500000, device=-1)
            arrs = ex.load_batch()
            names = ex.contig_names()
            ex.close()
            a = Oracle(0, 0, 500000, 0, contigs=names)
            a.batch(*arrs)
            prc, pb = 0, a.bed12()
        except RuntimeError:
            prc, pb = 1, None
        assert p.returncode == orc == prc, reg
        assert orc == 1 or p.stdout == ob == pb, reg


@pytest.mark.parametrize("block", range(3))
def test_feeder_on_damaged_files(block, tmp_path):
    """Truncated files, flipped bits, zeroed runs and wrong ISIZE trailers in fuzzed BAMs: what the unmodified reference prints
    before it stops (it checks neither CRC nor ISIZE: a block's length is zlib's total_out, bgzf.c:292-316; any failed or empty
    block ends the whole iteration, hts.c:1928-1963) must come out of the oracle and of the product's feeder (+ oracle walk)
    too — and nothing may crash.  (Round 1: this found that the feeder trusted the ISIZE trailer.)"""
    import random
    import fuzz_fixture as ff
    ref = os.path.join(ROOT, "oracle", "_ref", "regtools_ref")
    for seed in range(block * 12, block * 12 + 12):
        rnd = random.Random(seed)
        bam = ff.make_cigar_fuzz_bam(str(tmp_path / "a.bam"), seed)
        bad = str(tmp_path / "c.bam")
        mode = ff.damage_bam(bam, bad, seed)
        for reg in (".", "1:100-2000"):
            try:
                o = Oracle(0, 0, 500000, 0)
                o.extract_bam(bad, reg)
                want = (0, o.bed12())
            except RuntimeError:
                want = (1, "")
            try:
                ex = rt().JunctionsExtractor(bad, reg, 0, "XS", 0, 0, 500000, device=-1, n_threads=rnd.choice([1, 3]))
                arrs = ex.load_batch()
                names = ex.contig_names()
                ex.close()
                a = Oracle(0, 0, 500000, 0, contigs=names)
                a.batch(*arrs)
                got = (0, a.bed12())
            except RuntimeError:
                got = (1, "")
            assert got == want, (seed, mode, reg)
            if os.path.exists(ref):
                p = subprocess.run([ref, "junctions", "extract", "-s", "XS", "-a", "0", "-m", "0", "-r", reg, bad], capture_output=True, text=True)
                if p.returncode >= 0:                      # (the reference itself dies on a few of these)
                    assert (p.returncode, p.stdout) == want, (seed, mode, reg)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "regtools_ref")), reason="needs oracle/_ref/regtools_ref (dev container)")
def test_feeder_on_damaged_indexes(tmp_path):
    """Truncated, bit-flipped, zeroed and empty .bai files: the product's index loader and chunk planner must give what
    hts_idx_load / hts_itr_query give on the same damaged bytes — the same exit code ("Unable to open BAM/SAM index") or the
    same alignments — and never crash (ad hoc at the end of round 1: 180 runs, 0 differences)."""
    import random
    import fuzz_fixture as ff
    ref = os.path.join(ROOT, "oracle", "_ref", "regtools_ref")
    for seed in range(16):
        rnd = random.Random(seed)
        bam = ff.make_cigar_fuzz_bam(str(tmp_path / "i.bam"), seed)
        idx = bytearray(open(bam + ".bai", "rb").read())
        mode = rnd.choice(["trunc", "flip", "zero8", "empty"])
        if mode == "trunc":
            idx = idx[:rnd.randrange(0, len(idx))]
        elif mode == "flip":
            for _ in range(rnd.choice([1, 2])):
                idx[rnd.randrange(0, len(idx))] ^= 1 << rnd.randrange(8)
        elif mode == "zero8":
            i = rnd.randrange(0, max(1, len(idx) - 8))
            idx[i:i + 8] = bytes(8)
        else:
            idx = bytearray()
        open(bam + ".bai", "wb").write(idx)
        for reg in (".", "1:100-2000", "2"):
            p = subprocess.run([ref, "junctions", "extract", "-s", "XS", "-a", "0", "-m", "0", "-r", reg, bam], capture_output=True, text=True, timeout=60)
            try:
                ex = rt().JunctionsExtractor(bam, reg, 0, "XS", 0, 0, 500000, device=-1, n_threads=2)
                arrs = ex.load_batch()
                names = ex.contig_names()
                ex.close()
                a = Oracle(0, 0, 500000, 0, contigs=names)
                a.batch(*arrs)
                got = (0, a.bed12())
            except RuntimeError:
                got = (1, "")
            if p.returncode >= 0:
                assert (p.returncode, p.stdout) == got, (seed, mode, reg)


# ---- the equivalence the device feeder's `-r` path rests on -----------------------------------------------
def _endpos(pos, meta, off, cig):
    """bam_endpos (sam.c:336-342): pos + reference length of the CIGAR for a mapped alignment with a CIGAR, else pos + 1."""
    n = len(pos)
    ncig = (off[1:] - off[:-1]).astype(np.int64)
    op = cig & 0xF
    ln = (cig >> 4).astype(np.int64)
    ref = np.where(np.isin(op, [0, 2, 3, 7, 8]), ln, 0)                # M D N = X consume the reference
    csum = np.concatenate([[0], np.cumsum(ref)])
    rlen = csum[off[1:].astype(np.int64)] - csum[off[:-1].astype(np.int64)]
    unmapped = ((meta >> 16) & 4) != 0
    return pos.astype(np.int64) + np.where(unmapped | (ncig == 0), 1, rlen), n


@pytest.mark.parametrize("bam", [HCC, KAT, SYN, "generated"])
def test_region_iterator_equals_the_overlap_test_over_the_whole_file(bam, tmp_path):
    """hts_itr_next over the index's chunks (the host reader) returns exactly the alignments of the whole file with tid equal,
    pos < end and endpos > beg, in file order — the test cigar_scan applies when a large `-r` region streams through the
    device feeder as one byte span.  Random regions: whole contigs, windows of every size, empty stretches, odd bounds."""
    if bam == "generated":
        bam = str(tmp_path / "g.bam")
        subprocess.check_call([BAMGEN, "gen", "--out", bam, "--config", "c3", "--reads", "60000", "--seed", "21", "--level", "1"],
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    ex = rt().JunctionsExtractor(bam, ".", 0, "XS", 8, 70, 500000, device=-1, n_threads=2)
    tid, pos, meta, off, cig = ex.load_batch()
    names = ex.contig_names()
    ex.close()
    endpos, n = _endpos(pos, meta, off, cig)
    rng = np.random.default_rng(5)
    present = sorted(set(int(t) for t in np.unique(tid) if t >= 0))
    checked = nonempty = 0
    for k in range(60):
        t = int(rng.choice(present)) if k % 7 else int(rng.integers(0, len(names)))
        on = pos[tid == t]
        lo, hi = (int(on.min()), int(on.max())) if len(on) else (0, 1000)
        if k % 5 == 0:
            region, beg, end = names[t], 0, 2**31 - 1
        else:
            b = int(rng.integers(max(lo - 2000, 0), hi + 2000))
            w = int(rng.choice([1, 50, 2000, 100000, 30000000]))
            region, beg, end = f"{names[t]}:{b + 1}-{b + w}", b, b + w          # 1-based inclusive string -> 0-based [beg, end)
        want = np.flatnonzero((tid == t) & (pos < end) & (endpos > beg))
        rx = rt().JunctionsExtractor(bam, region, 0, "XS", 8, 70, 500000, device=-1, n_threads=2)
        rtid, rpos, rmeta, roff, rcig = rx.load_batch()
        rx.close()
        assert len(rtid) == len(want), (region, len(rtid), len(want))
        assert np.array_equal(rpos, pos[want]) and np.array_equal(rmeta, meta[want]), region
        assert np.array_equal(roff[1:] - roff[:-1], (off[1:] - off[:-1])[want]), region
        checked += 1; nonempty += int(len(want) > 0)
    assert checked == 60 and nonempty >= 20
