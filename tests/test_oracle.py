"""CPU: the oracle (oracle/jx_oracle.c) is pinned to the reference's goldens, its gtest known answers and
to outputs of the unmodified reference on the committed KAT fixtures (tests/golden/kat, make_golden.py)."""
import os

import numpy as np
import pytest

import synth
from oracle_py import REF_BIN, Oracle, ref_available, ref_extract

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
HCC = os.path.join(GOLD, "hcc1395", "test_hcc1395.bam")
MODES = {"XS": 0, "RF": 1, "FR": 2, "intron-motif": 3}


def run_oracle(bam, args, fasta=None, check=True):
    a, m, M, s, r, t = 8, 70, 500000, 0, ".", "XS"
    it = iter(args)
    for k in it:
        v = next(it)
        if k == "-a": a = int(v)
        elif k == "-m": m = int(v)
        elif k == "-M": M = int(v)
        elif k == "-s": s = MODES[v]
        elif k == "-r": r = v
        elif k == "-t": t = v
    o = Oracle(a & 0xFFFFFFFF, m & 0xFFFFFFFF, M & 0xFFFFFFFF, s, t, fasta=fasta)
    if check:
        o.extract_bam(bam, r)
    else:
        try:
            o.extract_bam(bam, r)
        except RuntimeError as e:
            o.failed = str(e)
    return o


@pytest.mark.parametrize("args,golden", [
    (["-s", "XS"], "expected-a.out"), (["-s", "XS", "-a", "30"], "expected-a30.out"),
    (["-s", "RF"], "expected-stranded-a.out"), (["-s", "RF", "-a", "30"], "expected-stranded-a30.out"),
    (["-s", "XS", "-m", "8039", "-M", "8039"], "expected-i8039-I8039.out"),
    (["-s", "XS", "-r", "1:22405013-22405020"], "expected-r1:22405013-22405020.out")])
def test_oracle_reference_goldens(args, golden):
    assert run_oracle(HCC, args).bed12() == open(os.path.join(GOLD, "hcc1395", golden)).read()


def manifest():
    rows = []
    for line in open(os.path.join(GOLD, "kat", "MANIFEST.tsv")):
        bam, out, args = line.rstrip("\n").split("\t")
        rows.append((bam, out, args.split()))
    return rows


@pytest.mark.parametrize("bam,out,args", [m for m in manifest() if not m[2][0] == "ctor"])
def test_oracle_matches_reference_outputs_on_kat(bam, out, args):
    assert run_oracle(os.path.join(GOLD, "kat", bam), args).bed12() == open(os.path.join(GOLD, "kat", out)).read()


@pytest.mark.parametrize("bam,out,args", [m for m in manifest() if m[2][0] == "ctor"])
def test_oracle_ctor_quirk(bam, out, args):
    """8-arg ctor: min_intron := min_anchor, get_all_junctions unfiltered (junctions_extractor.h:199-205, .cc:238-246)."""
    _, region, strandness, tag, anchor, _min_intron, max_intron = args
    o = Oracle(int(anchor), int(anchor), int(max_intron), int(strandness), tag)
    o.extract_bam(os.path.join(GOLD, "kat", bam), region)
    lines = []
    for j in o.table():
        lines.append("\t".join(map(str, [o.l.jxo_contig(o.h, int(j["tid"])).decode(), j["thick_start"], j["thick_end"],
                                         "JUNC%08d" % j["name_index"], j["read_count"], chr(j["strand"]), j["start"], j["end"],
                                         j["left_ok"], j["right_ok"]])) + "\n")
    assert "".join(lines) == open(os.path.join(GOLD, "kat", out)).read()


def test_oracle_gtest_add_junction():
    """tests/lib/junctions/test_junctions_extractor.cc:102-141."""
    o = Oracle(contigs=["chr1"])
    for a in [(10000, 10200, 9900, 10300, "+"), (10000, 10200, 9500, 10200, "+"), (10000, 10200, 9950, 10700, "+"),
              (8000, 8500, 7000, 10000, "+"), (8000, 8500, 7000, 10000, "-")]:
        o.add(0, *a)
    assert o.bed12() == ("chr1\t7000\t10000\tJUNC00000002\t1\t+\t7000\t10000\t255,0,0\t2\t1000,1500\t0,1500\n"
                         "chr1\t7000\t10000\tJUNC00000003\t1\t-\t7000\t10000\t255,0,0\t2\t1000,1500\t0,1500\n"
                         "chr1\t9500\t10700\tJUNC00000001\t3\t+\t9500\t10700\t255,0,0\t2\t500,500\t0,700\n")


def test_closed_form_equals_state_machine():
    """SURVEY Appendix A.2 (what the CUDA kernel computes) vs the literal state machine, on random CIGARs."""
    rng = np.random.default_rng(3)
    REFC, BRK, ANC = {0, 7, 2, 8, 3}, {3, 2, 8, 1, 4}, {0, 7}
    for _ in range(3000):
        n = int(rng.integers(2, 12))
        ops = [(int(rng.integers(0, 16)), int(rng.integers(0, 300))) for _ in range(n)]
        words = np.array([l << 4 | o for o, l in ops], np.uint32)
        pos = int(rng.integers(0, 1 << 20))
        o = Oracle(0, 0, 0xFFFFFFFF, 0, contigs=["c"])
        o.record_candidates()
        o.batch(np.array([0], np.int32), np.array([pos], np.int32), np.array([ord("+")], np.uint32),
                np.array([0, n], np.uint32), words)
        got = [(int(c["start"]), int(c["end"]), int(c["thick_start"]), int(c["thick_end"]), int(c["k"])) for c in o.candidates()]
        want = []
        for k, (op, ln) in enumerate(ops):
            if op != 3:
                continue
            start = pos + sum(l for (p, l) in ops[:k] if p in REFC)
            left = 0
            for p, l in reversed(ops[:k]):
                if p in BRK: break
                if p in ANC: left += l
            right = 0
            for p, l in ops[k + 1:]:
                if p in BRK: break
                if p in ANC: right += l
            m = 0xFFFFFFFF
            want.append((start & m, (start + ln) & m, (start - left) & m, (start + ln + right) & m, k))
        assert got == want, (ops, got, want)


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref/regtools_ref not built (needs /root/reference)")
@pytest.mark.parametrize("args", [["-s", "XS"], ["-s", "RF", "-a", "3"], ["-s", "XS", "-r", "10:1-700000"]])
def test_oracle_vs_live_reference_on_fresh_bam(args, tmp_path):
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bam = str(tmp_path / "fresh.bam")
    subprocess.check_call([os.path.join(root, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "30000",
                           "--seed", "2024", "--threads", "2"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    rc, out = ref_extract(bam, args)
    assert rc == 0
    assert run_oracle(bam, args).bed12() == out


from conftest import motif_manifest  # noqa: E402


@pytest.mark.parametrize("bam,fa,out,rc,args,err", motif_manifest())
def test_oracle_intron_motif_goldens(bam, fa, out, rc, args, err, motif_fastas):
    """FASTA given: strand from the intron motif, XS/flag only for '?', reverse-complement quirk of the reused
    Junction object, clipped fetches, missing contig -> runtime_error (junctions_extractor.cc:325-359,548-584)."""
    o = run_oracle(os.path.join(GOLD, "kat", bam), args, fasta=motif_fastas[fa], check=False)
    if rc:
        assert err and getattr(o, "failed", "").startswith(err)
    else:
        assert o.bed12() == open(os.path.join(GOLD, "motif", out)).read()


def test_fasta_fixtures_are_reproducible(motif_fastas):
    import hashlib
    got = {k: hashlib.sha256(open(v, "rb").read()).hexdigest() for k, v in motif_fastas.items()}
    want = dict(line.split() for line in open(os.path.join(GOLD, "motif", "FASTA.sha256")))
    assert got == want


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref/regtools_ref not built (needs /root/reference)")
@pytest.mark.parametrize("args", [["-s", "XS"], ["-s", "intron-motif", "-a", "3"], ["-s", "RF", "-r", "10:1-700000"]])
def test_oracle_vs_live_reference_with_fasta_on_fresh_bam(args, tmp_path, motif_fastas):
    """Differential run against the unmodified reference with a FASTA (intron-motif strand first) on a BAM neither has seen."""
    import subprocess
    from oracle_py import REF_BIN
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bam = str(tmp_path / "fresh.bam")
    subprocess.check_call([os.path.join(root, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "20000",
                           "--seed", "4242", "--threads", "2"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    fa = motif_fastas["synth"]                                   # contigs 1/10/2 of the `tiny` shape
    p = subprocess.run([REF_BIN, "junctions", "extract"] + args + [bam, fa], capture_output=True, text=True)
    assert p.returncode == 0
    o = run_oracle(bam, args, fasta=fa)
    assert o.bed12() == p.stdout
    assert len(set(l.split("\t")[5] for l in p.stdout.splitlines())) >= 2


def test_barcode_file_order_is_first_seen_insertion_replay():
    """F10 (`-b`, not built in the product): the reference's barcode file (tests/golden/barcodes, made by the unmodified
    reference) lists a junction's barcodes in std::unordered_map iteration order after one copy-assignment per read.
    That order equals inserting the junction's distinct barcodes in FIRST-SEEN order into one map (oracle/bc_replay.cc):
    checked here by re-deriving candidates and barcodes from the BAM with the test-side BAM reader."""
    import collections
    import struct
    import subprocess
    import zlib
    d = os.path.join(GOLD, "barcodes")
    raw, data, o = open(os.path.join(d, "bc.bam"), "rb").read(), b"", 0
    while o + 18 <= len(raw):
        bs = struct.unpack_from("<H", raw, o + 16)[0] + 1
        data += zlib.decompress(raw[o + 18:o + bs - 8], -15)
        o += bs
    l_text = struct.unpack_from("<i", data, 4)[0]
    p = 8 + l_text
    n_ref = struct.unpack_from("<i", data, p)[0]; p += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", data, p)[0]; p += 8 + l_name
    per = collections.OrderedDict()
    while p + 4 <= len(data):
        bl = struct.unpack_from("<i", data, p)[0]
        tid, pos, l_rn, mapq, bin_, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", data, p + 4)
        q = p + 36 + l_rn
        cig = struct.unpack_from(f"<{n_cig}I", data, q); q += 4 * n_cig + (l_seq + 1) // 2 + l_seq
        aux, strand, bc = data[q:p + 4 + bl], "?", "?"
        a = 0
        while a + 3 <= len(aux):                                   # only A and Z tags occur in this fixture
            tag, ty = aux[a:a + 2], aux[a + 2:a + 3]
            if ty == b"A":
                if tag == b"XS": strand = chr(aux[a + 3])
                a += 4
            else:
                e = aux.index(b"\0", a + 3)
                if tag == b"CB": bc = aux[a + 3:e].decode()
                a = e + 1
        cur = pos
        for w in cig:
            if w & 0xF == 3:
                dd = per.setdefault((cur, cur + (w >> 4), strand), collections.OrderedDict())
                dd[bc] = dd.get(bc, 0) + 1
            if w & 0xF in (0, 2, 3, 7, 8):
                cur += w >> 4
        p += 4 + bl
    lines = []
    for f in (l.split("\t") for l in open(os.path.join(d, "bc.bed")).read().splitlines()):
        left, right = (int(x) for x in f[10].split(","))
        v = per[(int(f[1]) + left, int(f[2]) - right, f[5])]
        assert sum(v.values()) == int(f[4])
        lines.append(" ".join(f"{b} {c}" for b, c in v.items()))
    oracle_dir = os.path.join(os.path.dirname(os.path.dirname(GOLD)), "oracle")
    exe = os.path.join(oracle_dir, "_ref", "bc_replay")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", oracle_dir, "-s"])
    got = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout
    assert got == open(os.path.join(d, "bc.barcodes")).read()
    assert max(len(v) for v in per.values()) > 40                 # several rehashes of the map


# ---- `-b` single-cell barcodes: the C oracle (+ bc_replay for the unordered_map order) against the reference ----------
BC2 = {"xs": ["-s", "XS"], "xs_a0": ["-s", "XS", "-a", "0"], "rf_m50": ["-s", "RF", "-m", "50"],
       "fr_region": ["-s", "FR", "-r", "10:1-60000"]}


def _bc_oracle(bam, args):
    from oracle_py import Oracle
    kw = dict(a=8, m=70, s=0, r=".")
    it = iter(args)
    for k in it:
        v = next(it)
        kw[k[1]] = {"XS": 0, "RF": 1, "FR": 2}[v] if k == "-s" else (v if k == "-r" else int(v))
    o = Oracle(kw["a"], kw["m"], 500000, kw["s"], barcodes=True)
    o.extract_bam(bam, kw["r"])
    return o


def test_barcode_oracle_matches_reference_goldens():
    """tests/golden/barcodes was written by the UNMODIFIED reference (make_golden.py --barcodes-only): BED12, the -b file
    and the number of `WARNING: No CB tag` lines must come out of the oracle byte for byte."""
    d = os.path.join(GOLD, "barcodes")
    o = _bc_oracle(os.path.join(d, "bc.bam"), ["-s", "XS"])
    assert o.bed12() == open(os.path.join(d, "bc.bed")).read()
    assert o.barcodes() == open(os.path.join(d, "bc.barcodes")).read()
    for tag, args in BC2.items():
        o = _bc_oracle(os.path.join(d, "bc2.bam"), args)
        assert o.bed12() == open(os.path.join(d, f"bc2.{tag}.bed")).read(), tag
        want = open(os.path.join(d, f"bc2.{tag}.barcodes")).read()
        assert o.barcodes() == want, tag
        assert o.barcodes_missing() == int(open(os.path.join(d, f"bc2.{tag}.warnings")).read()), tag
        assert len(want.splitlines()) == len(o.bed12().splitlines())
    big = max(int(l.split("\t")[0]) for l in open(os.path.join(d, "bc2.xs.barcodes")))
    assert big >= 150                                             # the hot junction: many rehashes of the reference's map


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="needs oracle/_ref/regtools_ref (dev container)")
@pytest.mark.parametrize("seed", [21, 22])
def test_barcode_oracle_live_differential(seed, tmp_path):
    """Fresh seeded fixtures through the unmodified reference and the oracle: -b file, BED12, warning count."""
    import subprocess
    import bc_fixture
    bam = bc_fixture.make_barcode_bam(str(tmp_path / "bc.bam"), seed=seed, n_reads=2500, n_barcodes=700, hot_barcodes=600)
    for args in (["-s", "XS"], ["-s", "FR", "-a", "3", "-m", "60"], ["-s", "XS", "-r", "2"]):
        p = subprocess.run([REF_BIN, "junctions", "extract"] + args + ["-b", str(tmp_path / "r.bc"), "-o", str(tmp_path / "r.bed"), bam],
                           capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        o = _bc_oracle(bam, args)
        assert o.bed12() == open(tmp_path / "r.bed").read()
        assert o.barcodes() == open(tmp_path / "r.bc").read()
        assert o.barcodes_missing() == p.stderr.count("WARNING: No CB tag found for alignment (id = 0)")


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="needs oracle/_ref/regtools_ref (dev container)")
def test_barcode_oracle_on_a_single_cell_shaped_bam(tmp_path):
    """bamgen --barcodes (the shape of the GPU at-scale test): 60k reads, 3000 skewed barcodes, a junction with > 800 of them."""
    import subprocess
    root = os.path.dirname(os.path.dirname(GOLD))
    bam = str(tmp_path / "sc.bam")
    subprocess.check_call([os.path.join(root, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "60000", "--seed", "9",
                           "--threads", "2", "--barcodes", "3000"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    p = subprocess.run([REF_BIN, "junctions", "extract", "-s", "XS", "-b", str(tmp_path / "r.bc"), "-o", str(tmp_path / "r.bed"), bam],
                       capture_output=True, text=True)
    assert p.returncode == 0
    o = _bc_oracle(bam, ["-s", "XS"])
    want = open(tmp_path / "r.bc").read()
    assert o.bed12() == open(tmp_path / "r.bed").read() and o.barcodes() == want
    assert o.barcodes_missing() == p.stderr.count("WARNING: No CB tag found for alignment (id = 0)") > 100
    assert max(int(l.split("\t")[0]) for l in want.splitlines()) > 800


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="needs oracle/_ref/regtools_ref (dev container)")
@pytest.mark.parametrize("block", range(3))
def test_barcode_differential_fuzz(block, tmp_path):
    """-b against the unmodified reference on fuzzed aux layouts (tests/fuzz_fixture.py: CB as Z or H, anywhere among other
    tags, odd barcode strings, reads without CB, XS of type Z, small BGZF blocks).  BED12, barcode file and warning count from
    the oracle; the product's host feeder must report the same number of untagged alignments and distinct barcodes
    (rtjx_load_barcodes, no GPU needed)."""
    import subprocess
    import fuzz_fixture as ff
    import regtools_b200 as rt
    for seed in range(block * 12, block * 12 + 12):
        bam = ff.make_barcode_fuzz_bam(str(tmp_path / "f.bam"), seed)
        for args in ff.BARCODE_FUZZ_ARGS:
            p = subprocess.run([REF_BIN, "junctions", "extract"] + args + ["-b", str(tmp_path / "r.bc"), "-o", str(tmp_path / "r.bed"), bam],
                               capture_output=True, text=True)
            assert p.returncode == 0, p.stderr[-300:]
            o = _bc_oracle(bam, args)
            assert o.bed12() == open(tmp_path / "r.bed").read(), (seed, args)
            assert o.barcodes() == open(tmp_path / "r.bc").read(), (seed, args)
            warnings = p.stderr.count("WARNING: No CB tag found for alignment (id = 0)")
            assert o.barcodes_missing() == warnings
            ex = rt.JunctionsExtractor(bam, args[args.index("-r") + 1] if "-r" in args else ".", 0, device=-1)
            ex.output_barcodes_file_ = os.devnull
            ids = ex.load_barcodes()
            n_bc, n_missing = ex.barcode_stats()
            names = ex.barcode_names()
            ex.close()
            assert n_missing == warnings and n_bc == len(set(names)) == len(names) and (len(ids) == 0 or ids.max() < max(n_bc, 1))


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="needs oracle/_ref/regtools_ref (dev container)")
@pytest.mark.parametrize("block", range(3))
def test_cigar_differential_fuzz(block, tmp_path):
    """Random CIGARs against the unmodified reference (tests/fuzz_fixture.py: every op incl. P / = / X / H, zero-length ops, N
    lengths on both QC bounds, up to 20 ops, flags incl. unmapped-with-CIGAR, XS of type A / C / absent / twice; tag NH as the
    strand tag, -a/-m/-M corners, a region).  The index comes from the reference's own htslib (oracle/_ref/ref_index)."""
    import subprocess
    import fuzz_fixture as ff
    root = os.path.dirname(os.path.dirname(GOLD))
    for seed in range(block * 10, block * 10 + 10):
        bam = ff.make_cigar_fuzz_bam(str(tmp_path / "f.bam"), seed)
        for args in ff.CIGAR_FUZZ_ARGS:
            rc, out = ref_extract(bam, args)
            o = run_oracle(bam, args)
            assert rc == 0 and o.bed12() == out, (seed, args)
        # the generator's own BAI writer must serve the same BAM (it once crashed on CIGARs that consume no reference)
        subprocess.check_call([os.path.join(root, "tools", "bamgen"), "index", bam], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        rc, out = ref_extract(bam, ["-s", "XS", "-r", "1:100-2000"])
        assert rc == 0 and out == run_oracle(bam, ["-s", "XS", "-r", "1:100-2000"]).bed12()


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="needs oracle/_ref/regtools_ref (dev container)")
@pytest.mark.parametrize("block", range(3))
def test_motif_differential_fuzz(block, tmp_path):
    """Intron-motif strand mode against the unmodified reference on fuzzed genomes (tests/fuzz_fixture.py): BED12 bytes, and
    where the FASTA lacks a contig the same runtime_error text and exit code 1."""
    import subprocess
    import fuzz_fixture as ff
    for seed in range(block * 10, block * 10 + 10):
        bam, fa = ff.make_motif_fuzz_case(str(tmp_path), seed)
        for args in ff.MOTIF_FUZZ_ARGS:
            p = subprocess.run([REF_BIN, "junctions", "extract"] + args + ["-o", str(tmp_path / "r.bed"), bam, fa], capture_output=True, text=True)
            o = run_oracle(bam, args, fasta=fa, check=False)
            failed = getattr(o, "failed", None)
            if failed:
                assert p.returncode == 1 and failed.strip() in p.stderr, (seed, args)
            else:
                assert p.returncode == 0 and o.bed12() == open(tmp_path / "r.bed").read(), (seed, args)


def ctor_rows(o):
    """get_all_junctions() of the oracle in the column layout of `regtools_ref ctor` (oracle/ref_main.cc)."""
    return "".join(f"{o.l.jxo_contig(o.h, int(j['tid'])).decode()}\t{j['thick_start']}\t{j['thick_end']}\tJUNC{j['name_index']:08d}\t"
                   f"{j['read_count']}\t{chr(j['strand'])}\t{j['start']}\t{j['end']}\t{j['left_ok']}\t{j['right_ok']}\n" for j in o.table())


def ctor_fuzz_queries(seed):
    import random
    rnd = random.Random(seed)
    out = []
    for _ in range(6):
        a = rnd.randrange(1, 4000)
        out.append((f"{rnd.choice(['1', '10', '2'])}:{a}-{a + rnd.choice([1, 50, 3000, 600000])}", rnd.choice([0, 1, 2]), rnd.choice([0, 1, 8, 30]),
                    rnd.choice([500000, 100, 4000000000])))
    return out


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="needs oracle/_ref/regtools_ref (dev container)")
@pytest.mark.parametrize("block", range(2))
def test_ctor_differential_fuzz(block, tmp_path):
    """The second caller's path (8-arg ctor + get_all_junctions, cis_splice_effects_identifier.cc:288-290; min_intron :=
    min_anchor, junctions_extractor.h:199-200) on fuzzed BAMs, regions and parameters, against `regtools_ref ctor`."""
    import subprocess
    import fuzz_fixture as ff
    for seed in range(block * 10, block * 10 + 10):
        bam = ff.make_cigar_fuzz_bam(str(tmp_path / "f.bam"), seed)
        for reg, s, anchor, M in ctor_fuzz_queries(seed):
            p = subprocess.run([REF_BIN, "ctor", bam, reg, str(s), "XS", str(anchor), "70", str(M)], capture_output=True, text=True)
            o = Oracle(anchor, anchor, M & 0xFFFFFFFF, s)
            o.extract_bam(bam, reg)
            assert p.returncode == 0 and ctor_rows(o) == p.stdout, (seed, reg, s, anchor, M)
