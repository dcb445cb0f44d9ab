"""GPU parity of `junctions annotate` through the C ABI (rtjx_annotate), the Python mirror and the CLI:
against the reference's own golden and files written by the UNMODIFIED reference (tests/golden/annotate), and against the
oracle (oracle/ja_oracle.cc) on fresh seeds.  junctions_annotator.cc:66-81,94-114,128-311,367-388; gtf_parser.cc.

(File name: sorts last on purpose — this row was built after the round's GPU budget was spent; its kernel source has only
been run through the host emulation harness of tests/test_annotate_cpu.py so far.)
"""
import os
import subprocess

import pytest

import ann_fixture

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "annotate")
ORACLE = os.path.join(ROOT, "oracle", "_ref", "ja_oracle")
CASES = [("hcc1395", ""), ("s1", ""), ("s1", ".S"), ("s2", ""), ("s2", ".S"), ("s3", ""), ("s3", ".S")]


@pytest.fixture(scope="module")
def gen_fasta(tmp_path_factory):
    return ann_fixture.write_fasta(str(tmp_path_factory.mktemp("annfa") / "ref.fa"))


@pytest.fixture(scope="module")
def oracle():
    if not os.path.exists(ORACLE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    return ORACLE


def _inputs(tag, gen_fasta):
    fa = os.path.join(GOLD, "hcc1395.fa") if tag == "hcc1395" else gen_fasta
    return os.path.join(GOLD, tag + ".bed"), fa, os.path.join(GOLD, tag + ".gtf")


def _annotate(bed, fa, gtf, out, single=False):
    import regtools_b200 as rt
    a = rt.JunctionsAnnotator(bed, fa, gtf)
    a.skip_single_exon_genes_ = not single
    a.output_file_ = str(out)
    n = a.annotate_all()
    return n, open(out).read()


@pytest.mark.parametrize("tag,suffix", CASES)
def test_reference_goldens(tag, suffix, gen_fasta, tmp_path):
    bed, fa, gtf = _inputs(tag, gen_fasta)
    n, got = _annotate(bed, fa, gtf, tmp_path / "o.tsv", single=bool(suffix))
    want = open(os.path.join(GOLD, f"{tag}{suffix}.expected.tsv")).read()
    assert got == want
    assert n == len(want.splitlines()) - 1


@pytest.mark.parametrize("seed", [21, 22, 23])
def test_fresh_cases_match_oracle(seed, oracle, gen_fasta, tmp_path):
    bed, _, gtf = ann_fixture.make_annotation_case(str(tmp_path / "case"), seed, header=seed == 22, crlf=seed == 23)
    for flags, single in (([], False), (["-S"], True)):
        subprocess.check_call([oracle] + flags + ["-o", str(tmp_path / "want.tsv"), bed, gen_fasta, gtf], stderr=subprocess.DEVNULL)
        _, got = _annotate(bed, gen_fasta, gtf, tmp_path / "got.tsv", single=single)
        assert got == open(tmp_path / "want.tsv").read()


def test_many_junctions_and_item_buffer_growth(gen_fasta, tmp_path, monkeypatch):
    """s1's junctions 200 times over (100k threads reserving item space concurrently), once with the default buffer and
    once with a 7-word buffer that forces the grow-and-rerun path."""
    bed, fa, gtf = _inputs("s1", gen_fasta)
    body = open(bed).read()
    big = tmp_path / "big.bed"
    big.write_text(body * 200)
    want = open(os.path.join(GOLD, "s1.expected.tsv")).read().splitlines(keepends=True)
    want = want[0] + "".join(want[1:]) * 200
    n, got = _annotate(str(big), fa, gtf, tmp_path / "a.tsv")
    assert got == want and n == 200 * len(body.splitlines())
    monkeypatch.setenv("RTJX_ANNOTATE_ITEMS", "7")
    _, got = _annotate(str(big), fa, gtf, tmp_path / "b.tsv")
    assert got == want


def test_error_paths_keep_the_reference_order(gen_fasta, tmp_path):
    import regtools_b200 as rt
    bed, fa, gtf = _inputs("s1", gen_fasta)
    lines = open(bed).read().splitlines()
    want_all = open(os.path.join(GOLD, "s1.expected.tsv")).read().splitlines(keepends=True)
    no10 = tmp_path / "no10.fa"
    no10.write_text("".join(">" + x for x in open(fa).read().split(">")[1:] if not x.startswith("10 ")))
    with pytest.raises(RuntimeError, match="Unable to extract FASTA sequence for position 10:"):
        _annotate(bed, str(no10), gtf, tmp_path / "a.tsv")
    first10 = next(i for i, l in enumerate(want_all[1:]) if l.startswith("10\t"))
    assert open(tmp_path / "a.tsv").read() == "".join(want_all[:1 + first10])
    bed6 = tmp_path / "bed6.bed"
    bed6.write_text("\n".join("\t".join(l.split("\t")[:6]) for l in lines[:5]) + "\n")
    with pytest.raises(RuntimeError, match="BED line not in BED12 format. start: "):
        _annotate(str(bed6), fa, gtf, tmp_path / "b.tsv")
    assert open(tmp_path / "b.tsv").read() == want_all[0]
    blank = tmp_path / "blank.bed"
    blank.write_text("\n".join(lines[:30] + [""] + lines[30:]) + "\n")
    n, got = _annotate(str(blank), fa, gtf, tmp_path / "c.tsv")
    assert n == 30 and got == "".join(want_all[:31])
    with pytest.raises(RuntimeError, match="Unable to open GTF file."):
        _annotate(bed, fa, str(tmp_path / "nope.gtf"), tmp_path / "d.tsv")
    assert not os.path.exists(tmp_path / "d.tsv")
    assert rt.junctions_annotate(["annotate", "-o", str(tmp_path / "e.tsv"), bed, fa]) == 1          # Error parsing inputs!(2)


def test_cli(gen_fasta, tmp_path):
    exe = os.path.join(ROOT, "regtools_b200", "regtools")
    bed, fa, gtf = _inputs("s2", gen_fasta)
    p = subprocess.run([exe, "junctions", "annotate", "-S", "-o", str(tmp_path / "o.tsv"), bed, fa, gtf], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    want = open(os.path.join(GOLD, "s2.S.expected.tsv")).read()
    assert open(tmp_path / "o.tsv").read() == want
    assert p.stderr.count("position = ") == 2 * (len(want.splitlines()) - 1)
    assert p.stderr.endswith(f"\nAnnotated {len(want.splitlines()) - 1} lines.\n") and "Skipping single exon genes." not in p.stderr
    p = subprocess.run([exe, "junctions", "annotate", bed, fa, gtf], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout == open(os.path.join(GOLD, "s2.expected.tsv")).read() and "Skipping single exon genes." in p.stderr
    assert subprocess.run([exe, "junctions", "annotate", bed, fa], capture_output=True).returncode == 1
    assert subprocess.run([exe, "junctions", "annotate", "-h"], capture_output=True).returncode == 0
