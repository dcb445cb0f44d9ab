"""GPU parity of the intron-motif strand mode (FASTA as second positional argument): the CUDA path against outputs of
the UNMODIFIED reference (tests/golden/motif, made by tests/golden/make_golden.py) and against the oracle on seeded
batches.  junctions_extractor.cc:325-359 (motif first, -s only for '?'), :548-584 (2-mer fetch, reverse-complement
quirk of the reused Junction object), faidx.c:341-415 (clipping), :553-555 (missing contig -> runtime_error)."""
import io
import os
import subprocess

import numpy as np
import pytest

import synth
from conftest import motif_manifest
from oracle_py import Oracle

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = {"XS": 0, "RF": 1, "FR": 2, "intron-motif": 3}


def _args(args):
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


@pytest.mark.parametrize("bam,fa,out,rc,args,err", motif_manifest())
def test_reference_motif_goldens(bam, fa, out, rc, args, err, motif_fastas, golden_dir):
    import regtools_b200 as rt
    k = _args(args)
    ex = rt.JunctionsExtractor(os.path.join(golden_dir, "kat", bam), k["r"], k["s"], k["t"], k["a"], k["m"], k["M"], motif_fastas[fa])
    if rc:
        with pytest.raises(RuntimeError) as e:
            ex.identify_junctions_from_BAM()
        assert err and str(e.value).startswith(err.split(":")[0] + ":")  # "... for position 10:" (which junction is reported first is not defined on the GPU)
        ex.close()
        return
    ex.identify_junctions_from_BAM()
    buf = io.StringIO()
    ex.print_all_junctions(buf)
    ex.close()
    assert buf.getvalue() == open(os.path.join(golden_dir, "motif", out)).read()


def _random_fasta(path, contigs, length, seed):
    rng = np.random.default_rng(seed)
    with open(path, "wb") as f:
        for c in contigs:
            s = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)[rng.choice(9, length, p=[.23, .23, .23, .23, .02, .02, .01, .01, .02])]
            f.write(b">" + c.encode() + b"\n")
            b = s.tobytes()
            for i in range(0, len(b), 80):
                f.write(b[i:i + 80] + b"\n")
    return path


@pytest.mark.parametrize("device_resident", [False, True])
@pytest.mark.parametrize("strandness", [0, 1, 3])
def test_random_batches_with_fasta_match_oracle(strandness, device_resident, tmp_path):
    """Random genome (so every motif class, lower case and N occur), contig 2 shorter than its junctions (clipping)."""
    import regtools_b200 as rt
    contigs = ["1", "10", "2"]
    fa = str(tmp_path / "rnd.fa")
    rng = np.random.default_rng(5)
    with open(fa, "wb") as f:
        for c, n in zip(contigs, (2_100_000, 2_100_000, 700_000)):
            s = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)[rng.choice(9, n, p=[.23, .23, .23, .23, .02, .02, .01, .01, .02])].tobytes()
            f.write(b">" + c.encode() + b" x\n" + b"\n".join(s[i:i + 61] for i in range(0, n, 61)) + b"\n")
    arrs = synth.random_batch(31 + strandness, 40000, spliced_frac=0.4)
    ex = rt.JunctionsExtractor(strandness=strandness, ref=fa)
    ex.set_contigs(contigs)
    if device_resident:
        dev = [torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).cuda() for x in arrs]
        ex.scan_batch(*dev, n_junction_ops=synth.count_n_ops(arrs[4]))
    else:
        h = len(arrs[0]) // 2
        off = arrs[3]
        ex.scan_batch(arrs[0][:h], arrs[1][:h], arrs[2][:h], off[:h + 1], arrs[4][:off[h]])
        ex.scan_batch(arrs[0][h:], arrs[1][h:], arrs[2][h:], (off[h:] - off[h]).astype(np.uint32), arrs[4][off[h]:], first_ordinal=h)
    got = ex.junction_table()
    buf = io.StringIO()
    ex.print_all_junctions(buf)
    ex.close()
    o = Oracle(8, 70, 500000, strandness, contigs=contigs, fasta=fa)
    o.batch(*arrs)
    want = o.table()
    assert o.error() is None
    assert len(got) == len(want)
    for f_ in ("tid", "start", "end", "thick_start", "thick_end", "read_count", "name_index", "strand", "left_ok", "right_ok"):
        assert np.array_equal(got[f_], want[f_]), f_
    assert buf.getvalue() == o.bed12()
    assert len(set(got["strand"].tolist())) >= 3


def test_device_feeder_and_cli_with_fasta(tmp_path, motif_fastas):
    """Whole-file run through the device feeder (BGZF inflate + record split on the GPU) and the CLI binary."""
    import regtools_b200 as rt
    bam = str(tmp_path / "gen.bam")
    subprocess.check_call([os.path.join(ROOT, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "500000", "--seed", "12"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    fa = motif_fastas["synth"]                                             # same contig names and lengths as the `tiny` config
    o = Oracle(8, 70, 500000, 3, fasta=fa)
    o.extract_bam(bam)
    want = o.bed12()
    ex = rt.JunctionsExtractor(bam, ".", 3, "XS", 8, 70, 500000, fa)
    ex.identify_junctions_from_BAM()
    buf = io.StringIO()
    ex.print_all_junctions(buf)
    st = ex.stats()
    ex.close()
    assert buf.getvalue() == want
    assert st["inflated_bytes"] > 0
    exe = os.path.join(ROOT, "regtools_b200", "regtools")
    out = tmp_path / "cli.bed"
    p = subprocess.run([exe, "junctions", "extract", "-s", "intron-motif", "-o", str(out), bam, fa], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert out.read_text() == want
    p = subprocess.run([exe, "junctions", "extract", "-s", "intron-motif", bam], capture_output=True, text=True)
    assert p.returncode == 1 and "requires a fasta file" in p.stderr
    p = subprocess.run([exe, "junctions", "extract", "-s", "XS", bam, motif_fastas["kat_no10"]], capture_output=True, text=True)
    assert p.returncode == 1 and "Unable to extract FASTA sequence for position 10:" in p.stderr
