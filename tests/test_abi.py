"""CPU: the C-ABI library loads and exports every symbol include/rtjx.h declares; host-only behaviour."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "rtjx.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(rtjx_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported():
    from regtools_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(_lib.lib, s), f"{s} declared in include/rtjx.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table and header disagree"


def test_struct_sizes_match_header():
    from regtools_b200 import _lib
    assert C.sizeof(_lib.Junction) == 40 and C.sizeof(_lib.Candidate) == 24
    p = _lib.Params()
    _lib.lib.rtjx_params_default(C.byref(p))
    assert p.struct_size == C.sizeof(_lib.Params)
    assert (p.min_anchor, p.min_intron, p.max_intron, p.region, p.strand_tag) == (8, 70, 500000, b".", b"XS")


def test_unsupported_modes_are_refused():
    """-b barcodes are built (single handle); what stays refused is -b combined with contig shards."""
    from regtools_b200 import _lib
    p = _lib.Params()
    _lib.lib.rtjx_params_default(C.byref(p))
    p.barcode_out = b"x"
    p.shard_world, p.shard_rank = 2, 0
    h = C.c_void_p()
    assert _lib.lib.rtjx_create(C.byref(p), C.byref(h)) == _lib.RTJX_E_UNSUPPORTED
    assert b"shards" in _lib.lib.rtjx_last_error(None)
    # a plain handle has no barcode table
    p = _lib.Params()
    _lib.lib.rtjx_params_default(C.byref(p))
    p.device = -1
    assert _lib.lib.rtjx_create(C.byref(p), C.byref(h)) == 0
    assert _lib.lib.rtjx_write_barcodes(h, 1) == _lib.RTJX_E_STATE
    assert _lib.lib.rtjx_barcode_stats(h, None, None) == _lib.RTJX_E_STATE
    _lib.lib.rtjx_destroy(h)


def test_barcode_mode_without_gpu_fails_loudly():
    """-b has no CPU path either: the feeder's dictionary is host work, the tables are not."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import regtools_b200 as rt
    ex = rt.JunctionsExtractor(os.path.join(ROOT, "tests", "golden", "barcodes", "bc.bam"), ".", 0)
    ex.output_barcodes_file_ = os.devnull
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.identify_junctions_from_BAM()
    ex.close()


def test_compute_without_gpu_fails_loudly():
    """No CPU fallback: on a box without a CUDA device every compute entry point returns RTJX_E_CUDA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import numpy as np
    import regtools_b200 as rt
    ex = rt.JunctionsExtractor(os.path.join(ROOT, "tests", "golden", "hcc1395", "test_hcc1395.bam"), ".", 0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.identify_junctions_from_BAM()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.add_junction(rt.Junction("chr1", 1, 100, 0, 120, "+"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ex.scan_batch(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.uint32), np.array([0, 1], np.uint32),
                      np.array([16], np.uint32))
    host = rt.JunctionsExtractor(device=-1)
    with pytest.raises(RuntimeError, match="host-only"):
        host.add_junction(rt.Junction("chr1", 1, 100, 0, 120, "+"))


def test_cli_option_handling_needs_no_gpu(tmp_path):
    """The C++ front-end (regtools_b200/regtools): usage / help / argument errors of both sub-commands have the reference's
    exit codes and texts (junctions_main.cc:45-107, junctions_extractor.cc:42-143, junctions_annotator.cc:405-456) and never
    touch the device; a compute request on a box without a GPU fails loudly with exit code 1."""
    import subprocess
    exe = os.path.join(ROOT, "regtools_b200", "regtools")
    run = lambda *a: subprocess.run([exe] + list(a), capture_output=True, text=True)
    p = run()
    assert p.returncode == 0 and "Usage:\t\tregtools <command> [options]" in p.stderr and "Version:\t1.0.0" in p.stderr
    p = run("junctions")
    assert p.returncode == 0 and "extract\t\tIdentify exon-exon junctions from alignments." in p.stdout and "annotate\tAnnotate the junctions." in p.stdout
    p = run("junctions", "extract", "-h")
    assert p.returncode == 0 and "Usage:\t\tregtools junctions extract [options] indexed_alignments.bam" in p.stderr
    p = run("junctions", "extract", "x.bam")
    assert p.returncode == 1 and "Please supply strandness mode" in p.stderr
    p = run("junctions", "annotate", "-h")
    assert p.returncode == 0 and "Usage:\t\tregtools junctions annotate [options] junctions.bed ref.fa annotations.gtf" in p.stderr
    assert "-S include single exon genes" in p.stderr
    p = run("junctions", "annotate", "a.bed", "b.fa")
    assert p.returncode == 1 and "Error parsing inputs!(2)" in p.stderr
    p = run("junctions", "annotate", "-x", "a.bed", "b.fa", "c.gtf")
    assert p.returncode == 1 and "Error parsing inputs!(1)" in p.stderr
    p = run("junctions", "annotate", "a.bed", "b.fa", str(tmp_path / "missing.gtf"))
    assert p.returncode == 1 and "Unable to open GTF file." in p.stderr and "Reference: b.fa" in p.stderr and "Skipping single exon genes." in p.stderr
    p = run("variants", "annotate")
    assert p.returncode == 1 and "not built here" in p.stderr
    import torch
    if not torch.cuda.is_available():
        gold = os.path.join(ROOT, "tests", "golden", "annotate")
        p = run("junctions", "annotate", "-o", str(tmp_path / "o.tsv"), os.path.join(gold, "hcc1395.bed"), os.path.join(gold, "hcc1395.fa"),
                os.path.join(gold, "hcc1395.gtf"))
        assert p.returncode == 1 and "no CPU fallback" in p.stderr and not os.path.exists(tmp_path / "o.tsv")
        p = run("junctions", "extract", "-s", "XS", os.path.join(ROOT, "tests", "golden", "hcc1395", "test_hcc1395.bam"))
        assert p.returncode == 1 and "no CPU fallback" in p.stderr and p.stdout == ""


# atoi() semantics of -a / -m / -M (junctions_extractor.cc:52-60), repeated and odd -s / -t / -r: the parameter echo on stderr
# carries the parsed values, the missing BAM ends the run before any device work
ODD_NUMERIC_ARGS = [("regtools_ref", "extract", ["-s", "XS"] + a + ["nonexist.bam"]) for a in (
    ["-a", "12junk"], ["-a", "-5"], ["-a", "99999999999"], ["-m", "x"], ["-M", "0x10"], ["-a", "1e3"], ["-a", "+7"], ["-a", " 8"],
    ["-m", "4294967296"], ["-M", "-1"], ["-a", "007"], ["-s", "RF"], ["-s", "0"], ["-s", "3"], ["-t", "X"], ["-t", "LONGTAG"],
    ["-r", "1:1-2", "-r", "2"])]


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "regtools_ref_annotate")), reason="needs oracle/_ref (dev container)")
def test_cli_texts_equal_the_reference(tmp_path):
    """Exit code, stdout and stderr (after our three banner lines, which the reference's real CLI prints too but the test
    drivers in oracle/ do not) of the option / error paths that need no device, byte for byte against the unmodified
    reference classes."""
    import subprocess
    exe = os.path.join(ROOT, "regtools_b200", "regtools")
    bam = os.path.join(ROOT, "tests", "golden", "hcc1395", "test_hcc1395.bam")
    cases = [("regtools_ref", "extract", a) for a in (["-h"], ["x.bam"], ["-s", "XS"], ["-s", "bogus", bam], ["-s", "XS", "nonexist.bam"],
                                                       ["-s", "intron-motif", bam], ["-q", "-s", "XS", bam], ["-s", "XS", "-o", "o", "-r", "1:2-3", "-t", "ZS", "-b", "b", "nonexist.bam"])]
    cases += ODD_NUMERIC_ARGS
    cases += [("regtools_ref_annotate", "annotate", a) for a in (["-h"], ["a.bed", "b.fa"], ["-x", "a.bed", "b.fa", "c.gtf"], ["a.bed", "b.fa", "/nonexistent.gtf"],
                                                                 ["-S", "-o", "o.tsv", "a.bed", "b.fa", "/nonexistent.gtf"], ["a", "b", "c", "d"])]
    for ref_bin, sub, args in cases:
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", ref_bin), "junctions", sub] + args, capture_output=True, cwd=tmp_path)
        o = subprocess.run([exe, "junctions", sub] + args, capture_output=True, cwd=tmp_path)
        assert o.stderr.startswith(b"\nProgram:\tregtools\nVersion:\t1.0.0\n")
        ours_err = b"\n".join(o.stderr.split(b"\n")[3:])
        assert (r.returncode, r.stdout, r.stderr) == (o.returncode, o.stdout, ours_err), (sub, args)


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "regtools_ref_annotate")), reason="needs oracle/_ref (dev container)")
def test_python_mirror_texts_equal_the_reference(tmp_path):
    """regtools_b200.junctions_extract / junctions_annotate on the same option / error paths: exit code, stdout, stderr."""
    import subprocess
    import sys
    bam = os.path.join(ROOT, "tests", "golden", "hcc1395", "test_hcc1395.bam")
    runner = "import sys, regtools_b200 as rt\nsys.exit(getattr(rt, 'junctions_' + sys.argv[1])([sys.argv[1]] + sys.argv[2:]))"
    cases = [("regtools_ref", "extract", a) for a in (["-h"], ["x.bam"], ["-s", "XS"], ["-s", "bogus", bam], ["-s", "XS", "nonexist.bam"],
                                                       ["-s", "intron-motif", bam], ["-q", "-s", "XS", bam], ["-s", "XS", "-o"],
                                                       ["-s", "XS", "-o", "o", "-r", "1:2-3", "-t", "ZS", "-b", "b", "-a", "3", "-m", "4", "-M", "5", "nonexist.bam"])]
    cases += ODD_NUMERIC_ARGS
    cases += [("regtools_ref_annotate", "annotate", a) for a in (["-h"], ["a.bed", "b.fa"], ["-x", "a.bed", "b.fa", "c.gtf"], ["a.bed", "b.fa", "/nonexistent.gtf"],
                                                                 ["-S", "-o", "o.tsv", "a.bed", "b.fa", "/nonexistent.gtf"], ["-o"], ["a", "b", "c", "d"])]
    env = dict(os.environ, PYTHONPATH=ROOT)
    for ref_bin, sub, args in cases:
        r = subprocess.run([os.path.join(ROOT, "oracle", "_ref", ref_bin), "junctions", sub] + args, capture_output=True, cwd=tmp_path)
        o = subprocess.run([sys.executable, "-c", runner, sub] + args, capture_output=True, cwd=tmp_path, env=env)
        assert (r.returncode, r.stdout, r.stderr) == (o.returncode, o.stdout, o.stderr), (sub, args)


def test_reference_callers_compile_against_the_shim_header(tmp_path):
    """The drop-in claim of INTEGRATION.md §2, checked: the reference's OWN callers of the class — `junctions_main.cc`
    (CLI glue), `cis_splice_effects_identifier.cc` (the 8-arg-ctor caller) and its gtest file
    `tests/lib/junctions/test_junctions_extractor.cc` — compile, unchanged, with the shim's junctions_extractor.h in front of
    the reference's.  Needs the reference tree (dev container); nothing is linked or run."""
    import shutil
    import subprocess
    ref = "/root/reference"
    if not os.path.isdir(os.path.join(ref, "src", "junctions")) or shutil.which("g++") is None:
        pytest.skip("needs /root/reference and g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # a maintainer replaces src/junctions/junctions_extractor.h in place; here: a shadow of that directory (links to the
    # reference's files) in which only that header is the shim — quoted includes resolve next to the including file
    shadow = tmp_path / "junctions"
    shadow.mkdir()
    for f in os.listdir(os.path.join(ref, "src", "junctions")):
        if f != "junctions_extractor.h":
            os.symlink(os.path.join(ref, "src", "junctions", f), shadow / f)
    (shadow / "junctions_extractor.h").write_text('#include "%s"\n' % os.path.join(root, "regtools_b200", "csrc", "junctions_extractor.h"))
    inc = [f"-I{shadow}", f"-I{root}/include"] + [f"-I{ref}/{d}" for d in (
        "src/utils", "src/utils/htslib", "src/utils/bedtools/bedFile", "src/utils/bedtools/lineFileUtilities",
        "src/utils/bedtools/gzstream", "src/utils/bedtools/fileType", "src/utils/bedtools/stringUtilities", "src/gtf", "src/variants",
        "src/cis-splice-effects", "src/utils/gtest-1.7.0/include")]
    for src in (str(shadow / "junctions_main.cc"), os.path.join(ref, "src/cis-splice-effects/cis_splice_effects_identifier.cc"),
                os.path.join(ref, "tests/lib/junctions/test_junctions_extractor.cc")):
        p = subprocess.run(["g++", "-std=c++11", "-w", "-fsyntax-only", "-DRTJX_SHIM_CHECK", src] + inc, capture_output=True, text=True)
        assert p.returncode == 0, f"{src} does not compile against the shim:\n{p.stderr[:3000]}"
