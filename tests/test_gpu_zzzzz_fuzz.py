"""GPU parity on fuzzed BAMs (tests/fuzz_fixture.py): the CUDA path through the Python mirror against the oracle, which the
CPU suite pins to the unmodified reference on the same generators (tests/test_oracle.py::test_cigar_differential_fuzz,
::test_barcode_differential_fuzz).  Written at the end of round 1 without a GPU: sorts last in the suite."""
import io
import os

import pytest

import fuzz_fixture as ff
from oracle_py import Oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
MODES = {"XS": 0, "RF": 1, "FR": 2}


def _kw(args):
    kw = dict(a=8, m=70, M=500000, s=0, r=".", t="XS")
    it = iter(args)
    for k in it:
        v = next(it)
        if k in ("-a", "-m", "-M"):
            kw[k[1]] = int(v) & 0xFFFFFFFF
        elif k == "-s":
            kw["s"] = MODES[v]
        else:
            kw[k[1]] = v
    return kw


@pytest.mark.parametrize("block", range(3))
def test_fuzzed_cigars_match_oracle(block, tmp_path):
    import regtools_b200 as rt
    for seed in range(block * 10, block * 10 + 10):
        bam = ff.make_cigar_fuzz_bam(str(tmp_path / "f.bam"), seed)
        for args in ff.CIGAR_FUZZ_ARGS:
            k = _kw(args)
            ex = rt.JunctionsExtractor(bam, k["r"], k["s"], k["t"], k["a"], k["m"], k["M"])
            ex.identify_junctions_from_BAM()
            buf = io.StringIO()
            ex.print_all_junctions(buf)
            ex.close()
            o = Oracle(k["a"], k["m"], k["M"], k["s"], k["t"])
            o.extract_bam(bam, k["r"])
            assert buf.getvalue() == o.bed12(), (seed, args)


@pytest.mark.parametrize("block", range(3))
def test_fuzzed_barcodes_match_oracle(block, tmp_path, capfd):
    import regtools_b200 as rt
    for seed in range(block * 12, block * 12 + 12):
        bam = ff.make_barcode_fuzz_bam(str(tmp_path / "f.bam"), seed)
        for args in ff.BARCODE_FUZZ_ARGS:
            k = _kw(args)
            ex = rt.JunctionsExtractor(bam, k["r"], k["s"], k["t"], k["a"], k["m"], k["M"])
            ex.output_barcodes_file_ = os.devnull
            ex.identify_junctions_from_BAM()
            bed, bc = io.StringIO(), io.StringIO()
            ex.print_barcodes(bc)
            ex.print_all_junctions(bed)
            n_missing = ex.barcode_stats()[1]
            ex.close()
            o = Oracle(k["a"], k["m"], k["M"], k["s"], k["t"], barcodes=True)
            o.extract_bam(bam, k["r"])
            assert bed.getvalue() == o.bed12(), (seed, args)
            assert bc.getvalue() == o.barcodes(), (seed, args)
            assert n_missing == o.barcodes_missing()
    capfd.readouterr()


@pytest.mark.parametrize("block", range(3))
def test_fuzzed_motif_genomes_match_oracle(block, tmp_path):
    """Intron-motif strand mode on fuzzed genomes; a FASTA without contig 10 must raise the reference's text (which junction
    is named first is not defined on the GPU, so only the prefix up to the contig is compared)."""
    import regtools_b200 as rt
    for seed in range(block * 10, block * 10 + 10):
        bam, fa = ff.make_motif_fuzz_case(str(tmp_path), seed)
        for args in ff.MOTIF_FUZZ_ARGS:
            k = _kw(args) if args[1] != "intron-motif" else dict(_kw(["-s", "FR"] + args[2:]), s=3)
            o = Oracle(k["a"], k["m"], k["M"], k["s"], k["t"], fasta=fa)
            try:
                o.extract_bam(bam, k["r"])
                failed = None
            except RuntimeError as e:
                failed = str(e)
            ex = rt.JunctionsExtractor(bam, k["r"], k["s"], k["t"], k["a"], k["m"], k["M"], fa)
            if failed:
                with pytest.raises(RuntimeError) as err:
                    ex.identify_junctions_from_BAM()
                assert str(err.value).startswith(failed.split(":")[0] + ":"), (seed, args)
                ex.close()
                continue
            ex.identify_junctions_from_BAM()
            buf = io.StringIO()
            ex.print_all_junctions(buf)
            ex.close()
            assert buf.getvalue() == o.bed12(), (seed, args)


@pytest.mark.parametrize("block", range(2))
def test_fuzzed_variant_regions_match_oracle(block, tmp_path):
    """rtjx_run_regions (all regions of a BAM in one pass) on fuzzed BAMs and regions: region i must equal what the oracle's
    8-arg-ctor path (pinned to `regtools_ref ctor` by tests/test_oracle.py::test_ctor_differential_fuzz) gives for it."""
    import numpy as np
    import regtools_b200 as rt
    from test_oracle import ctor_fuzz_queries
    for seed in range(block * 10, block * 10 + 10):
        bam = ff.make_cigar_fuzz_bam(str(tmp_path / "f.bam"), seed)
        queries = ctor_fuzz_queries(seed)
        for s, anchor, M in sorted({(q[1], q[2], q[3]) for q in queries}):
            regs = [q[0] for q in queries if (q[1], q[2], q[3]) == (s, anchor, M)]
            ex = rt.JunctionsExtractor.from_region(bam, ".", s, "XS", anchor, 70, M & 0xFFFFFFFF)
            tables = ex.identify_junctions_in_regions(regs)
            ex.close()
            for reg, got in zip(regs, tables):
                o = Oracle(anchor, anchor, M & 0xFFFFFFFF, s)
                o.extract_bam(bam, reg)
                want = o.table()
                assert len(got) == len(want), (seed, reg)
                for f in ("tid", "start", "end", "thick_start", "thick_end", "read_count", "name_index", "strand", "left_ok", "right_ok"):
                    assert np.array_equal(got[f], want[f]), (seed, reg, f)


@pytest.mark.parametrize("block", range(2))
def test_damaged_files_match_oracle(block, tmp_path):
    """Truncated / bit-flipped / zeroed / wrong-ISIZE copies of fuzzed BAMs through the whole product: the BED12 must be what
    the oracle prints (= what the unmodified reference prints before it stops, tests/test_host_logic.py::
    test_feeder_on_damaged_files).  inflate_mode 0 / 1 take the host feeder on files this small; inflate_mode 2 forces the
    device feeder (lane-per-stream or warp-per-block inflate, record starts found on the device, chain walk), which has to
    decline — inflate status, record checks, a missed seed, an untrusted ISIZE — or agree."""
    import regtools_b200 as rt
    for seed in range(block * 12, block * 12 + 12):
        bam = ff.make_cigar_fuzz_bam(str(tmp_path / "a.bam"), seed)
        bad = str(tmp_path / "c.bam")
        mode = ff.damage_bam(bam, bad, seed)
        for reg in (".", "1:100-2000"):
            o = Oracle(0, 0, 500000, 0)
            try:
                o.extract_bam(bad, reg)
                want = o.bed12()
            except RuntimeError:
                want = None
            # (inflate_mode, RTJX_INFLATE_VARIANT): 2 forces the device feeder onto the damaged file — with the warp-per-block
            # decoder (1) and with the lane-per-stream decoder + match resolve (3) — and it must decline or agree
            for inflate_mode, variant in ((0, None), (1, None), (2, "1"), (2, "3")):
                if variant is None:
                    os.environ.pop("RTJX_INFLATE_VARIANT", None)
                else:
                    os.environ["RTJX_INFLATE_VARIANT"] = variant
                try:
                    ex = rt.JunctionsExtractor(bad, reg, 0, "XS", 0, 0, 500000, inflate_mode=inflate_mode)
                    try:
                        ex.identify_junctions_from_BAM()
                        buf = io.StringIO()
                        ex.print_all_junctions(buf)
                        got = buf.getvalue()
                    except RuntimeError:
                        got = None
                    ex.close()
                finally:
                    os.environ.pop("RTJX_INFLATE_VARIANT", None)
                assert got == want, (seed, mode, reg, inflate_mode, variant)


@pytest.mark.parametrize("block", range(2))
def test_fuzzed_annotation_inputs_match_oracle(block, tmp_path):
    """`junctions annotate` on the pathological GTF / BED generators of tests/test_annotate_fuzz.py (mixed-strand transcripts,
    60-exon transcripts with equal starts, malformed intervals, CRLF, mid-file headers, missing contigs): exit status and the
    bytes on disk of the CUDA path against the oracle, which the CPU suite pins to the unmodified reference on the same seeds."""
    import random
    import subprocess
    import ann_fixture
    import regtools_b200 as rt
    from test_annotate_fuzz import _case_mixed, _case_ties
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    oracle = os.path.join(root, "oracle", "_ref", "ja_oracle")
    if not os.path.exists(oracle):
        subprocess.check_call(["make", "-C", os.path.join(root, "oracle"), "-s"])
    fa = ann_fixture.write_fasta(str(tmp_path / "ref.fa"), [("1", 60000), ("2", 30000)])
    for seed in range(block * 20, block * 20 + 20):
        for gtf_text, bed_text in (_case_mixed(random.Random(seed), 0.05), _case_ties(random.Random(10_000 + seed))):
            gtf, bed = tmp_path / "a.gtf", tmp_path / "j.bed"
            gtf.write_text(gtf_text)
            with open(bed, "w", newline="") as f:
                f.write(bed_text)
            for single in (False, True):
                want_path, got_path = tmp_path / "want.tsv", tmp_path / "got.tsv"
                for p in (want_path, got_path):
                    if p.exists():
                        p.unlink()
                rc = subprocess.run([oracle] + (["-S"] if single else []) + ["-o", str(want_path), str(bed), fa, str(gtf)],
                                    capture_output=True).returncode
                a = rt.JunctionsAnnotator(str(bed), fa, str(gtf))
                a.skip_single_exon_genes_ = not single
                a.output_file_ = str(got_path)
                try:
                    a.annotate_all()
                    got_rc = 0
                except RuntimeError:
                    got_rc = 1
                assert got_rc == rc, (seed, single)
                assert (got_path.read_text() if got_path.exists() else None) == (want_path.read_text() if want_path.exists() else None), (seed, single)
