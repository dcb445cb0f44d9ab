"""CPU: `junctions annotate`.

* the oracle (oracle/ja_oracle.cc) is pinned to the reference's own golden and to files written by the UNMODIFIED reference
  (tests/golden/annotate, tests/golden/make_golden.py --annotate-only), and — in the dev container — to live runs of
  oracle/_ref/regtools_ref_annotate on fresh seeds;
* the PRODUCT's kernel source (regtools_b200/csrc/annotate.cu) and the host code around it (annotate.cc) are run through the
  host emulation harness tests/emul (g++ build of the kernel body, malloc/memcpy stand-ins for the CUDA runtime) and must
  give the same bytes: that checks their logic where no GPU exists; the GPU build is tested in test_gpu_zzz_annotate.py;
* the shipped library refuses to annotate without a CUDA device (no CPU fallback).
"""
import os
import subprocess

import pytest

import ann_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "annotate")
ORACLE = os.path.join(ROOT, "oracle", "_ref", "ja_oracle")
REF = os.path.join(ROOT, "oracle", "_ref", "regtools_ref_annotate")
EMUL = os.path.join(ROOT, "build", "emul", "annotate_emul")
CASES = [("hcc1395", ""), ("s1", ""), ("s1", ".S"), ("s2", ""), ("s2", ".S"), ("s3", ""), ("s3", ".S")]


@pytest.fixture(scope="session")
def tools():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s"])
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "emul"), "-s"])
    return {"oracle": ORACLE, "emul": EMUL}


@pytest.fixture(scope="session")
def gen_fasta(tmp_path_factory):
    return ann_fixture.write_fasta(str(tmp_path_factory.mktemp("annfa") / "ref.fa"))


def _inputs(tag, gen_fasta):
    fa = os.path.join(GOLD, "hcc1395.fa") if tag == "hcc1395" else gen_fasta
    return os.path.join(GOLD, tag + ".bed"), fa, os.path.join(GOLD, tag + ".gtf")


def _run(exe, flags, bed, fa, gtf, out, env=None):
    p = subprocess.run([exe] + flags + ["-o", str(out), bed, fa, gtf], capture_output=True, text=True, env=env)
    return p.returncode, (open(out).read() if os.path.exists(out) else None), p.stderr


@pytest.mark.parametrize("which", ["oracle", "emul"])
@pytest.mark.parametrize("tag,suffix", CASES)
def test_reference_goldens(which, tag, suffix, tools, gen_fasta, tmp_path):
    bed, fa, gtf = _inputs(tag, gen_fasta)
    rc, out, err = _run(tools[which], ["-S"] if suffix else [], bed, fa, gtf, tmp_path / "o.tsv")
    assert rc == 0, err
    want = open(os.path.join(GOLD, f"{tag}{suffix}.expected.tsv")).read()
    assert out == want
    assert err.endswith(f"\nAnnotated {len(want.splitlines()) - 1} lines.\n")
    assert err.count("position = ") == 2 * (len(want.splitlines()) - 1)


def test_goldens_cover_the_branches():
    rows = [l.split("\t") for l in open(os.path.join(GOLD, "s1.expected.tsv")).read().splitlines()[1:]]
    assert {r[10] for r in rows} == {"DA", "NDA", "D", "A", "N"}
    assert {r[5] for r in rows} >= {"+", "-", "?"}
    assert sum(r[8] != "0" for r in rows) > 20 and sum("," in r[16] for r in rows) > 50 and sum("," in r[14] for r in rows) > 10
    assert any(len(r[6]) < 5 for r in rows)                       # clipped fetch at the end of a sequence
    assert open(os.path.join(GOLD, "s1.expected.tsv")).read() != open(os.path.join(GOLD, "s1.S.expected.tsv")).read()


def test_item_buffer_grows_and_reruns(tools, gen_fasta, tmp_path):
    """A 7-word item buffer overflows on the first launch; the rerun with the exact size must give the same bytes."""
    bed, fa, gtf = _inputs("s1", gen_fasta)
    rc, out, err = _run(tools["emul"], [], bed, fa, gtf, tmp_path / "o.tsv", env=dict(os.environ, RTJX_ANNOTATE_ITEMS="7"))
    assert rc == 0, err
    assert out == open(os.path.join(GOLD, "s1.expected.tsv")).read()


def _error_inputs(gen_fasta, tmp_path):
    bed, fa, gtf = _inputs("s1", gen_fasta)
    lines = open(bed).read().splitlines()
    gl = open(gtf).read().splitlines()
    no10 = tmp_path / "no10.fa"
    no10.write_text("".join(">" + x for x in open(fa).read().split(">")[1:] if not x.startswith("10 ")))
    bad_bed = tmp_path / "bad.bed"
    bad_bed.write_text("\n".join(lines[:50] + ["\t".join(lines[50].split("\t")[:6])] + lines[51:]) + "\n")
    bed6 = tmp_path / "bed6.bed"
    bed6.write_text("\n".join("\t".join(l.split("\t")[:6]) for l in lines[:5]) + "\n")
    blank = tmp_path / "blank.bed"
    blank.write_text("\n".join(lines[:30] + [""] + lines[30:]) + "\n")
    bad_gtf = tmp_path / "bad.gtf"
    bad_gtf.write_text("\n".join(gl[:10] + ["\t".join(gl[10].split("\t")[:8])] + gl[11:]) + "\n")
    return {"missing_contig": (bed, str(no10), gtf), "differing_fields": (str(bad_bed), fa, gtf), "bed6": (str(bed6), fa, gtf),
            "blank_line": (str(blank), fa, gtf), "gtf_8_fields": (bed, fa, str(bad_gtf)), "no_gtf": (bed, fa, str(tmp_path / "nope.gtf"))}


@pytest.mark.parametrize("which", ["oracle", "emul"])
def test_error_paths_keep_the_reference_order(which, tools, gen_fasta, tmp_path):
    """What is on disk when the reference stops: nothing for a GTF problem, the lines before the offending junction else."""
    cases = _error_inputs(gen_fasta, tmp_path)
    want_all = open(os.path.join(GOLD, "s1.expected.tsv")).read().splitlines(keepends=True)
    rc, out, err = _run(tools[which], [], *cases["blank_line"], tmp_path / "a.tsv")
    assert rc == 0 and out == "".join(want_all[:31]) and "Annotated 30 lines." in err       # BED_BLANK ends the loop, exit 0
    rc, out, err = _run(tools[which], [], *cases["differing_fields"], tmp_path / "b.tsv")
    assert rc == 1 and out == "".join(want_all[:51]) and "Differing number of BED fields" in err
    rc, out, err = _run(tools[which], [], *cases["bed6"], tmp_path / "c.tsv")
    assert rc == 1 and out == want_all[0] and "BED line not in BED12 format. start: " in err
    rc, out, err = _run(tools[which], [], *cases["missing_contig"], tmp_path / "d.tsv")
    first10 = next(i for i, l in enumerate(want_all[1:]) if l.startswith("10\t"))
    assert rc == 1 and out == "".join(want_all[:1 + first10]) and "Unable to extract FASTA sequence for position 10:" in err
    for k, msg in (("gtf_8_fields", "Expected 9 fields in GTF line."), ("no_gtf", "Unable to open GTF file.")):
        rc, out, err = _run(tools[which], [], *cases[k], tmp_path / (k + ".tsv"))
        assert rc == 1 and out is None and msg in err                                       # the output is never opened


@pytest.mark.skipif(not os.path.exists(REF), reason="needs oracle/_ref/regtools_ref_annotate (dev container)")
@pytest.mark.parametrize("seed", [11, 12, 13])
def test_live_differential_against_the_unmodified_reference(seed, tools, gen_fasta, tmp_path):
    bed, _, gtf = ann_fixture.make_annotation_case(str(tmp_path / "case"), seed, header=seed == 12, crlf=seed == 13)
    os.remove(str(tmp_path / "case" / "ref.fa"))
    for flags in ([], ["-S"]):
        p = subprocess.run([REF, "junctions", "annotate"] + flags + ["-o", str(tmp_path / "r.tsv"), bed, gen_fasta, gtf], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        want = open(tmp_path / "r.tsv").read()
        chatter = p.stderr[p.stderr.index("position = "):]
        for which in ("oracle", "emul"):
            rc, out, err = _run(tools[which], flags, bed, gen_fasta, gtf, tmp_path / "o.tsv")
            assert rc == 0 and out == want, which
            if which == "emul":
                assert err == chatter                          # the product reproduces the reference's stderr lines too
    # error paths, reference against both
    for name, (b, f, g) in _error_inputs(gen_fasta, tmp_path).items():
        if name in ("differing_fields",):
            continue                                           # the reference exit()s there without flushing its ofstream
        ref_out = tmp_path / ("ref_" + name + ".tsv")
        p = subprocess.run([REF, "junctions", "annotate", "-o", str(ref_out), b, f, g], capture_output=True, text=True)
        for which in ("oracle", "emul"):
            rc, out, err = _run(tools[which], [], b, f, g, tmp_path / (which + name + ".tsv"))
            assert rc == p.returncode, (name, which)
            assert out == (open(ref_out).read() if os.path.exists(ref_out) else None), (name, which)


def test_shipped_library_has_no_cpu_path(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import regtools_b200 as rt
    a = rt.JunctionsAnnotator(os.path.join(GOLD, "hcc1395.bed"), os.path.join(GOLD, "hcc1395.fa"), os.path.join(GOLD, "hcc1395.gtf"))
    a.output_file_ = str(tmp_path / "o.tsv")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        a.annotate_all()
    assert rt.junctions_annotate(["annotate", "-o", str(tmp_path / "o.tsv"), os.path.join(GOLD, "hcc1395.bed"),
                                  os.path.join(GOLD, "hcc1395.fa"), os.path.join(GOLD, "hcc1395.gtf")]) == 1
    with pytest.raises(RuntimeError, match=r"Error parsing inputs!\(2\)"):
        rt.JunctionsAnnotator().parse_options(["annotate", "only.bed"])
    with pytest.raises(rt.CmdlineHelpException):
        rt.JunctionsAnnotator().parse_options(["annotate", "-h"])
    # tests/lib/junctions/test_junctions_annotator.cc:35-45 (ParseInput)
    ja1 = rt.JunctionsAnnotator()
    assert ja1.parse_options(["annotate", "test.bed", "test.fa", "test.gtf"]) == 0 and ja1.gtf_file() == "test.gtf"


def test_gtest_known_answers_of_the_gtf_parser(tools, tmp_path):
    """tests/lib/gtf/test_gtf_parser.cc:86-121: the EP300 exon 22:12791-14103 files its transcript under bin 37359 (level 0 of
    the reference's offsets, 32678 + 4681 + 0), gene EP300 / ENSG00000100393.  Seen from outside: a junction in that bin that
    ends exactly on the exon's start finds the transcript, on either harness."""
    gtf = tmp_path / "one.gtf"
    attr = ('ccds_id "CCDS14010"; exon_id "ENSE00001343011"; exon_number "1"; gene_biotype "protein_coding"; gene_id "ENSG00000100393"; '
            'gene_name "EP300"; gene_source "ensembl_havana"; p_id "P5137"; tag "CCDS"; transcript_id "ENST00000263253"; '
            'transcript_name "EP300-001"; transcript_source "ensembl_havana"; tss_id "TSS138009"')
    gtf.write_text("22\tprotein_coding\texon\t12791\t14103\t.\t+\t.\t" + attr + "\n"
                   "22\tprotein_coding\texon\t38192\t38300\t.\t+\t.\t" + attr.replace('"1"', '"2"') + "\n")
    bed = tmp_path / "j.bed"
    # known junction 14103 -> 38192 (the first line of the reference's golden), and one ending on the first exon's start
    bed.write_text("22\t14006\t38288\tJ1\t38\t+\t14006\t38288\t255,0,0\t2\t97,97\t0,24185\n"
                   "22\t12000\t12890\tJ2\t1\t+\t12000\t12890\t255,0,0\t2\t50,100\t0,790\n")
    for which in ("oracle", "emul"):
        rc, out, err = _run(tools[which], [], str(bed), os.path.join(GOLD, "hcc1395.fa"), str(gtf), tmp_path / (which + ".tsv"))
        assert rc == 0, err
        rows = [l.split("\t") for l in out.splitlines()[1:]]
        assert rows[0][:3] == ["22", "14103", "38192"] and rows[0][10:] == ["DA", "1", "1", "1", "EP300", "ENSG00000100393", "ENST00000263253"]
        assert rows[1][:3] == ["22", "12050", "12791"] and rows[1][10:] == ["A", "0", "1", "0", "EP300", "ENSG00000100393", "ENST00000263253"]


def test_side_bench_tool_runs_with_the_emulation_harness(tools, tmp_path):
    """tools/bench_annotate.py (the measurement script of this row) end to end on a small workload, with the harness standing
    in for the GPU: the JSON line parses and its parity field says the sample matched the CPU arm byte for byte."""
    import json
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_annotate.py"), "--impl", "emul", "--contigs", "1", "--mb", "3",
                        "--junctions", "3000", "--sample", "800", "--steps", "1", "--warmup", "0", "--dir", str(tmp_path)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["parity"] == "byte-identical on the sample" and line["value"] > 0 and line["cpu_baseline"]["value"] > 0
    assert line["unit"] == "junctions/s" and line["impl"] == "emul"


def test_junctions_from_stdin(tools, gen_fasta, tmp_path):
    """BedFile reads "stdin" / "-" (bedFile.cpp:99-101): `regtools junctions extract ... | regtools junctions annotate - ref.fa x.gtf`."""
    bed, fa, gtf = _inputs("s2", gen_fasta)
    for name in ("-", "stdin"):
        p = subprocess.run([tools["emul"], "-o", str(tmp_path / "o.tsv"), name, fa, gtf], stdin=open(bed), capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
        assert open(tmp_path / "o.tsv").read() == open(os.path.join(GOLD, "s2.expected.tsv")).read()
