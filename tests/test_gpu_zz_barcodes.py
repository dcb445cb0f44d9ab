"""GPU parity of the `-b` single-cell mode (F10: set_junction_barcode junctions_extractor.cc:362-374, the barcode branch of
add_junction :203-215, print_barcodes junctions_extractor.h:99-111 driven by print_all_junctions :255-257,272-273).

The CUDA path keeps a (junction, barcode) table, folds it into the junction table on the device and prints each junction's
barcodes by replaying them into a std::unordered_map in first-seen order.  Checked against files written by the UNMODIFIED
reference (tests/golden/barcodes, tests/golden/make_golden.py --barcodes-only) and against the oracle on fresh fixtures.

(File name: this module sorts last on purpose — the mode was written after the round's GPU budget was spent, so it is the
one part of the suite that first runs on a GPU at round end.)
"""
import io
import os
import subprocess

import numpy as np
import pytest

import bc_fixture
from oracle_py import Oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "barcodes")
MODES = {"XS": 0, "RF": 1, "FR": 2}
BC2 = {"xs": ["-s", "XS"], "xs_a0": ["-s", "XS", "-a", "0"], "rf_m50": ["-s", "RF", "-m", "50"],
       "fr_region": ["-s", "FR", "-r", "10:1-60000"]}


def _kw(args):
    kw = dict(a=8, m=70, s=0, r=".")
    it = iter(args)
    for k in it:
        v = next(it)
        kw[k[1]] = MODES[v] if k == "-s" else (v if k == "-r" else int(v))
    return kw


def _product(bam, args, tmp_path, fasta="NA", **opts):
    """BED12 text, barcode file text, (n_barcodes, n_missing) of one `-b` run through the Python mirror."""
    import regtools_b200 as rt
    k = _kw(args)
    ex = rt.JunctionsExtractor(bam, k["r"], k["s"], "XS", k["a"], k["m"], 500000, fasta, **opts)
    ex.output_barcodes_file_ = str(tmp_path / "out.barcodes")
    ex.output_file_ = str(tmp_path / "out.bed")
    ex.identify_junctions_from_BAM()
    ex.print_all_junctions()
    stats = ex.barcode_stats()
    buf = io.StringIO()
    ex.print_barcodes(buf)                                   # same bytes through a pipe
    table = ex.junction_table()
    ex.close()
    bc = open(tmp_path / "out.barcodes").read()
    assert buf.getvalue() == bc
    return open(tmp_path / "out.bed").read(), bc, stats, table


def _oracle(bam, args, fasta=None):
    k = _kw(args)
    o = Oracle(k["a"], k["m"], 500000, k["s"], barcodes=True, fasta=fasta)
    o.extract_bam(bam, k["r"])
    return o


def test_reference_golden_small(tmp_path):
    bed, bc, (n_bc, n_missing), _ = _product(os.path.join(GOLD, "bc.bam"), ["-s", "XS"], tmp_path)
    assert bed == open(os.path.join(GOLD, "bc.bed")).read()
    assert bc == open(os.path.join(GOLD, "bc.barcodes")).read()
    assert (n_bc, n_missing) == (59, 24)                     # 58 barcodes + "?"; 24 WARNING lines of the reference


@pytest.mark.parametrize("tag", sorted(BC2))
def test_reference_goldens_bc2(tag, tmp_path, capfd):
    """3 contigs, a hot junction with 166 barcodes, proxy-2 strands, QC- and anchor-filtered junctions, CB behind Z/B tags."""
    bed, bc, (_, n_missing), table = _product(os.path.join(GOLD, "bc2.bam"), BC2[tag], tmp_path)
    assert bed == open(os.path.join(GOLD, f"bc2.{tag}.bed")).read()
    assert bc == open(os.path.join(GOLD, f"bc2.{tag}.barcodes")).read()
    want_warn = int(open(os.path.join(GOLD, f"bc2.{tag}.warnings")).read())
    assert n_missing == want_warn
    assert capfd.readouterr().err.count("WARNING: No CB tag found for alignment (id = 0)") == want_warn
    # the junction-level table (fold of the pair table) equals the oracle's, field by field
    want = _oracle(os.path.join(GOLD, "bc2.bam"), BC2[tag]).table()
    assert len(table) == len(want)
    for f in ("tid", "start", "end", "thick_start", "thick_end", "read_count", "name_index", "strand", "left_ok", "right_ok"):
        assert np.array_equal(table[f], want[f]), f


@pytest.mark.parametrize("opts", [dict(), dict(batch_reads=1024, table_log2=8), dict(n_threads=1, batch_reads=4096)])
@pytest.mark.parametrize("seed", [31, 32])
def test_fresh_fixtures_match_oracle(seed, opts, tmp_path):
    """Seeded fixtures the goldens do not cover; small batches and a 256-slot table force batch boundaries inside a
    junction's barcode list and several rehashes of the (junction, barcode) table."""
    bam = bc_fixture.make_barcode_bam(str(tmp_path / "f.bam"), seed=seed, n_reads=6000, n_barcodes=900, hot_barcodes=800,
                                      missing=0.1)
    for args in (["-s", "XS"], ["-s", "FR", "-a", "3", "-m", "60"], ["-s", "XS", "-r", "2"]):
        bed, bc, (_, n_missing), _ = _product(bam, args, tmp_path, **opts)
        o = _oracle(bam, args)
        assert bed == o.bed12(), args
        assert bc == o.barcodes(), args
        assert n_missing == o.barcodes_missing()
        assert sum(int(l.split("\t")[4]) for l in bed.splitlines()) == \
            sum(int(x.rsplit(":", 1)[1]) for l in bc.splitlines() for x in l.split("\t")[1].split(","))


def test_barcodes_with_intron_motif_strands(tmp_path):
    """FASTA + -b: the strand comes from the motif (kernel instantiation MOTIF + BC), the barcode id must survive it.
    Contig 1 = (GTAG)n gives '+' or '?' by phase, contig 10 = (CTAC)n gives '-' or '?', contig 2 is shorter than its
    junctions (clipped fetch -> '?', then XS decides)."""
    fa = str(tmp_path / "g.fa")
    with open(fa, "w") as f:
        for name, unit, n in (("1", "GTAG", 400000), ("10", "CTAC", 400000), ("2", "ACGT", 20000)):
            s = unit * (n // 4)
            f.write(f">{name}\n" + "\n".join(s[i:i + 80] for i in range(0, n, 80)) + "\n")
    bam = bc_fixture.make_barcode_bam(str(tmp_path / "f.bam"), seed=41, n_reads=4000)
    for args in (["-s", "XS"], ["-s", "RF"]):
        bed, bc, _, _ = _product(bam, args, tmp_path, fasta=fa)
        o = _oracle(bam, args, fasta=fa)
        assert o.error() is None
        assert bed == o.bed12() and bc == o.barcodes()
        assert {"+", "-"} <= set(l.split("\t")[5] for l in bed.splitlines())


def test_untagged_bam_at_scale(tmp_path):
    """A generated BAM without CB tags: every junction's line is `1<TAB>?:<read_count>`, one warning per n_cigar > 1
    alignment; 300k reads exercise the fold with the contention of the hot junctions."""
    import regtools_b200 as rt
    bam = str(tmp_path / "big.bam")
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "300000",
                           "--seed", "5"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    plain = rt.JunctionsExtractor(bam, ".", 0)
    plain.identify_junctions_from_BAM()
    want_tab = plain.junction_table()
    buf = io.StringIO()
    plain.print_all_junctions(buf)
    plain.close()
    ex = rt.JunctionsExtractor(bam, ".", 0)
    ex.output_barcodes_file_ = str(tmp_path / "o.bc")
    import contextlib
    with open(os.devnull, "w") as dn, contextlib.redirect_stderr(dn):      # hundreds of thousands of WARNING lines
        ex.identify_junctions_from_BAM()
    out = io.StringIO()
    ex.print_all_junctions(out)
    got_tab = ex.junction_table()
    n_bc, n_missing = ex.barcode_stats()
    ex.close()
    assert out.getvalue() == buf.getvalue() and len(buf.getvalue().splitlines()) > 100
    assert got_tab.tobytes() == want_tab.tobytes()             # -b does not change the junction table (same feeder order)
    lines = open(tmp_path / "o.bc").read().splitlines()
    assert lines == ["1\t?:" + l.split("\t")[4] for l in buf.getvalue().splitlines()]
    from oracle_py import Oracle as _O
    o = _O(8, 70, 500000, 0, barcodes=True)
    o.extract_bam(bam)
    assert n_bc == 1 and n_missing == o.barcodes_missing() > 10000


def test_single_cell_shape_at_scale(tmp_path):
    """bamgen --barcodes: 300k reads of the `tiny` shape (3 contigs, Zipf junction weights) with CB:Z drawn from 5000 skewed
    barcodes on 97 % of the reads: ~42k printed (junction, barcode) pairs, hot junctions with thousands of distinct barcodes (the
    replayed unordered_map rehashes a dozen times), several device batches, table growth from 2^10 slots."""
    bam = str(tmp_path / "sc.bam")
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "300000",
                           "--seed", "17", "--barcodes", "5000"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    import contextlib
    with open(os.devnull, "w") as dn, contextlib.redirect_stderr(dn):
        bed, bc, (n_bc, n_missing), _ = _product(bam, ["-s", "XS"], tmp_path, batch_reads=65536, table_log2=10)
    o = _oracle(bam, ["-s", "XS"])
    assert bed == o.bed12()
    assert bc == o.barcodes()
    assert n_missing == o.barcodes_missing() > 500 and 1000 < n_bc <= 5001
    assert max(int(l.split("\t")[0]) for l in bc.splitlines()) > 2000          # 3067 distinct barcodes on the hottest junction


@pytest.mark.parametrize("device_resident", [False, True])
@pytest.mark.parametrize("strandness", [0, 1])
def test_batch_level_barcodes_match_oracle(strandness, device_resident):
    """rtjx_scan_batch with a `bc` column (ids from rtjx_intern_barcode) on random batches, split over several calls so a
    junction's barcode list crosses batch boundaries; host-resident and device-resident input."""
    import synth
    import regtools_b200 as rt
    arrs = synth.random_batch(11 + strandness, 40000, spliced_frac=0.3)
    tid, pos, meta, off, cigar = arrs
    rng = np.random.default_rng(5)
    names = ["BC%05d-1" % i for i in range(700)] + ["?"]
    bc_name_idx = np.minimum((len(names) * rng.random(len(tid)) ** 2).astype(np.int64), len(names) - 1)      # skewed
    ex = rt.JunctionsExtractor(strandness=strandness)
    ex.output_barcodes_file_ = os.devnull
    ex.set_contigs(["1", "10", "2"])
    ids = np.array([ex.intern_barcode(n) for n in reversed(names)], dtype=np.uint32)[::-1]              # ids != name index
    bc = ids[bc_name_idx].astype(np.uint32)
    bounds = np.linspace(0, len(tid), 4).astype(int)
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        o = off[lo:hi + 1].astype(np.uint32)
        c = cigar[o[0]:o[-1]]
        sub = [tid[lo:hi], pos[lo:hi], meta[lo:hi], (o - o[0]).astype(np.uint32), c, bc[lo:hi]]
        if device_resident:
            sub = [torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).cuda() for x in sub]
        ex.scan_batch(*sub[:5], first_ordinal=int(lo), n_junction_ops=synth.count_n_ops(c), bc=sub[5])
        if device_resident:
            torch.cuda.synchronize()
    bed, bcs = io.StringIO(), io.StringIO()
    ex.print_barcodes(bcs)
    ex.print_all_junctions(bed)                                # (also writes the barcode lines to os.devnull)
    ex.close()
    orc = Oracle(8, 70, 500000, strandness, contigs=["1", "10", "2"], barcodes=True)
    orc.batch_barcodes(tid, pos, meta, off, cigar, bc_name_idx.astype(np.uint32), names)
    assert bed.getvalue() == orc.bed12()
    assert bcs.getvalue() == orc.barcodes()
    assert max(int(l.split("\t")[0]) for l in bcs.getvalue().splitlines()) > 20


def test_cli_writes_the_barcode_file(tmp_path):
    exe = os.path.join(ROOT, "regtools_b200", "regtools")
    bam = os.path.join(GOLD, "bc2.bam")
    p = subprocess.run([exe, "junctions", "extract", "-s", "XS", "-a", "0", "-b", str(tmp_path / "c.bc"), "-o", str(tmp_path / "c.bed"), bam],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert open(tmp_path / "c.bed").read() == open(os.path.join(GOLD, "bc2.xs_a0.bed")).read()
    assert open(tmp_path / "c.bc").read() == open(os.path.join(GOLD, "bc2.xs_a0.barcodes")).read()
    assert p.stderr.count("WARNING: No CB tag found for alignment (id = 0)") == int(open(os.path.join(GOLD, "bc2.xs_a0.warnings")).read())
    assert "Barcode file: " in p.stderr
    # BED12 on stdout, barcodes to the file
    p = subprocess.run([exe, "junctions", "extract", "-s", "XS", "-b", str(tmp_path / "d.bc"), bam], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout == open(os.path.join(GOLD, "bc2.xs.bed")).read()
    assert open(tmp_path / "d.bc").read() == open(os.path.join(GOLD, "bc2.xs.barcodes")).read()


def test_rerun_and_refusals(tmp_path):
    """A -b handle is fed by rtjx_run only; a second run on the same handle starts from an empty dictionary and table."""
    import regtools_b200 as rt
    bam = os.path.join(GOLD, "bc2.bam")
    ex = rt.JunctionsExtractor(bam, ".", 0)
    ex.output_barcodes_file_ = str(tmp_path / "r.bc")
    for _ in range(2):
        ex.clear()
        ex.identify_junctions_from_BAM()
        buf = io.StringIO()
        ex.print_barcodes(buf)
        assert buf.getvalue() == open(os.path.join(GOLD, "bc2.xs.barcodes")).read()
    with pytest.raises(RuntimeError, match="carries no barcodes"):
        ex.add_junction(rt.Junction("1", 100, 300, 50, 350, "+"))
    with pytest.raises(RuntimeError, match="needs rtjx_batch.bc"):
        ex.scan_batch(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.uint32), np.array([0, 1], np.uint32),
                      np.array([16], np.uint32))
    with pytest.raises(RuntimeError, match="no barcode mode"):
        ex.identify_junctions_in_regions(["1:1-1000"])
    ex.close()
