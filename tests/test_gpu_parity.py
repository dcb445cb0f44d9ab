"""GPU parity tests: the CUDA path (through the C ABI) vs the CPU oracle and the reference's goldens.

Bit-exact bar: every field of every junction (coordinates, counts, first-seen name index, printed
strand, anchor flags) and the BED12 text must be identical.
"""
import io
import os

import numpy as np
import pytest

import synth
from oracle_py import Oracle

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


def _rt():
    import regtools_b200 as rt
    return rt


def tables_equal(gpu, cpu):
    assert len(gpu) == len(cpu), f"junction count {len(gpu)} vs oracle {len(cpu)}"
    for f in ("tid", "start", "end", "thick_start", "thick_end", "read_count", "name_index", "strand", "left_ok", "right_ok"):
        bad = np.nonzero(gpu[f] != cpu[f])[0]
        assert bad.size == 0, f"field {f} differs at {bad[:5]}: gpu {gpu[f][bad[:5]]} oracle {cpu[f][bad[:5]]}"


def run_gpu_batch(arrs, strandness=0, a=8, m=70, M=500000, contigs=("1", "10", "2"), device_resident=False, split=1,
                  variant=0, cfg=0, known=True):
    """variant 0/5 = block-per-tile cigar_scan (default), 8 = warp-pipelined cigar_scan; both feed junction_merge."""
    rt = _rt()
    ex = rt.JunctionsExtractor(strandness=strandness, min_anchor_length=a, min_intron_length=m, max_intron_length=M,
                               scan_variant=variant, scan_cfg=cfg)
    ex.set_contigs(list(contigs))
    tid, pos, meta, off, cigar = arrs
    n = len(tid)
    bounds = np.linspace(0, n, split + 1).astype(int)
    for i in range(split):
        lo, hi = bounds[i], bounds[i + 1]
        if hi == lo:
            continue
        o = off[lo:hi + 1].astype(np.uint32)
        c = cigar[o[0]:o[-1]]
        o = (o - o[0]).astype(np.uint32)
        sub = (tid[lo:hi], pos[lo:hi], meta[lo:hi], o, c)
        if device_resident:
            dev = [torch.from_numpy(np.ascontiguousarray(x).view(np.int32)).cuda() for x in sub]
            if dev[4].numel() == 0:
                dev[4] = torch.zeros(4, dtype=torch.int32, device="cuda")[:0]
            ex.scan_batch(*dev, first_ordinal=int(lo), n_junction_ops=synth.count_n_ops(c))
            torch.cuda.synchronize()
        else:
            ex.scan_batch(*sub, first_ordinal=int(lo), n_junction_ops=synth.count_n_ops(c) if known else 0)
    table = ex.junction_table()
    buf = io.StringIO()
    ex.print_all_junctions(buf)
    stats = ex.stats()
    ex.close()
    return table, buf.getvalue(), stats


def run_oracle_batch(arrs, strandness=0, a=8, m=70, M=500000, contigs=("1", "10", "2")):
    o = Oracle(a, m, M, strandness, "XS", contigs=list(contigs))
    o.batch(*arrs)
    return o.table(), o.bed12()


@pytest.mark.parametrize("variant,known", [(0, True), (8, True), (0, False), (8, False)])
@pytest.mark.parametrize("strandness", [0, 1, 2])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_batches_match_oracle(seed, strandness, variant, known):
    arrs = synth.random_batch(seed, 20000)
    g_tab, g_bed, st = run_gpu_batch(arrs, strandness, variant=variant, known=known)
    assert st["candidates"] == synth.count_n_ops(arrs[4]) - 1      # read 3 has tid -1: its N op is never emitted
    o_tab, o_bed = run_oracle_batch(arrs, strandness)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed


@pytest.mark.parametrize("device_resident", [False, True])
def test_split_batches_and_device_resident(device_resident):
    arrs = synth.random_batch(11, 50000, spliced_frac=0.3)
    g_tab, g_bed, st = run_gpu_batch(arrs, 0, device_resident=device_resident, split=7)
    o_tab, o_bed = run_oracle_batch(arrs, 0)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed
    assert st["kernel_launches"] >= 7 and st["reads"] == 50000


@pytest.mark.parametrize("cfg", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("n_reads", [1, 511, 512, 513, 4099, 70001])
def test_fused_scan_tile_configs_and_ragged_sizes(cfg, n_reads):
    """Every tile configuration of the fused kernel, batch sizes around the tile boundaries, dense splicing
    (alignments outside the staged slab window read their ops from global memory)."""
    arrs = synth.random_batch(100 + n_reads, n_reads, spliced_frac=0.6 if n_reads < 5000 else 0.2)
    g_tab, g_bed, _ = run_gpu_batch(arrs, 0, variant=6, cfg=cfg)
    o_tab, o_bed = run_oracle_batch(arrs, 0)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed


def test_fused_scan_large_batch_matches_two_kernel_path():
    """300k reads, 1200 junctions: fused (6) and two-kernel (5) paths against the oracle on a device-resident batch."""
    arrs = synth.random_batch(9, 300000, spliced_frac=0.1, catalog_per_contig=400)
    big = tuple(arrs)
    g_tab, g_bed, _ = run_gpu_batch(big, 0, variant=6, device_resident=True, split=1)
    v5_tab, v5_bed, _ = run_gpu_batch(big, 0, variant=5, device_resident=True, split=1)
    o_tab, o_bed = run_oracle_batch(big, 0)
    tables_equal(g_tab, o_tab)
    tables_equal(v5_tab, o_tab)
    assert g_bed == o_bed == v5_bed


@pytest.mark.parametrize("a,m,M", [(0, 0, 0xFFFFFFFF), (30, 70, 500000), (8, 8039, 8039), (1, 1, 69)])
def test_parameter_corners(a, m, M):
    arrs = synth.random_batch(5, 15000, spliced_frac=0.4)
    g_tab, g_bed, _ = run_gpu_batch(arrs, 0, a, m, M)
    o_tab, o_bed = run_oracle_batch(arrs, 0, a, m, M)
    tables_equal(g_tab, o_tab)
    assert g_bed == o_bed


def test_known_answer_vectors():
    """SURVEY 8a KAT table (verified against the reference binary)."""
    C = synth.cig
    p, m = ord("+"), ord("-")
    reads = [
        (0, 1000, 0, 60, p, C("50M100N50M")), (0, 1000, 0, 0, p, C("10S40M100N50M")),
        (0, 1000, 0, 60, p, C("5H50M100N50M5H")), (0, 1000, 0, 60, p, C("20M2X28M100N50M")),
        (0, 1000, 0, 60, p, C("20=30M100N25=25M")), (0, 1000, 0, 60, p, C("50M100N20M1I29M")),
        (0, 1000, 0, 60, p, C("50M100N20M3D30M")),
        (0, 2000, 0, 60, p, C("50M100N200N50M")), (0, 3000, 0, 60, m, C("50M100N30M200N20M")),
        (0, 4000, 0, 60, p, C("50M100N")), (0, 5000, 0, 60, p, C("50M69N50M")), (0, 5500, 0, 60, p, C("50M70N50M")),
        (0, 6000, 0, 60, p, C("7M100N50M")), (0, 6000, 0, 60, p, C("50M100N7M")),
        (0, 6957, 0, 60, p, C("50M100N7M")), (0, 7000, 0, 60, p, C("7M100N50M")),
        (0, 8000, 0, 60, ord("."), C("50M100N50M")), (0, 8000, 0, 60, 0, C("50M100N50M")),
        (0, 9500, 0, 60, p, C("50M3P100N50M")), (0, 10000, 4, 0, p, C("50M100N50M")),
        (0, 10000, 256 | 512 | 1024 | 2048, 0, p, C("50M100N50M")),
        (0, 11000, 0, 60, p, C("50M500001N50M")), (0, 11000, 0, 60, p, C("50M500000N50M")),
        (1, 100, 0, 60, p, C("50M100N50M")), (2, 100, 0, 60, p, C("50M100N50M")),
    ]
    arrs = synth.batch_from_reads(reads)
    for a in (8, 0):
        g_tab, g_bed, _ = run_gpu_batch(arrs, 0, a=a)
        o_tab, o_bed = run_oracle_batch(arrs, 0, a=a)
        tables_equal(g_tab, o_tab)
        assert g_bed == o_bed
    assert "1\t1000\t1200\tJUNC00000001\t6\t+\t1000\t1200\t255,0,0\t2\t50,50\t0,150\n" in g_bed
    assert "1\t1000\t1190\tJUNC00000002\t1\t+\t1000\t1190\t255,0,0\t2\t40,50\t0,140\n" in g_bed


def test_hot_junction_and_long_cigars():
    """One junction supported by 200k reads (atomic contention) + reads with thousands of ops."""
    C = synth.cig
    reads = [(0, 5000 - (i % 90) - 1, 99, 60, ord("+"), C(f"{(i % 90) + 1}M1000N{100 - (i % 90)}M")) for i in range(200000)]
    long_ops = []
    for k in range(3000):
        long_ops += C("5M80N")
    reads.append((0, 900000, 0, 60, ord("-"), long_ops + C("5M")))
    reads.append((1, 10, 0, 60, ord("-"), C("10M") * 40000 + C("70N10M")))
    arrs = synth.batch_from_reads(reads)
    o_tab, o_bed = run_oracle_batch(arrs, 0)
    for variant in (6, 5):
        g_tab, g_bed, _ = run_gpu_batch(arrs, 0, variant=variant)
        tables_equal(g_tab, o_tab)
        assert g_bed == o_bed
        assert g_tab["read_count"].max() == 200000


def test_many_unique_junctions_grow_table():
    rng = np.random.default_rng(7)
    n = 300000
    starts = np.sort(rng.integers(0, 2_000_000_000, n)).astype(np.int64)
    lens = rng.integers(70, 5000, n)
    reads = [(0, int(s), 0, 60, ord("+"), [50 << 4, int(l) << 4 | 3, 50 << 4]) for s, l in zip(starts, lens)]
    arrs = synth.batch_from_reads(reads)
    rt = _rt()
    ex = rt.JunctionsExtractor(strandness=0, table_log2=12)
    ex.set_contigs(["1"])
    # a tiny first batch creates the 2^12-slot table; the rest forces a rehash into a larger one
    tid, pos, meta, off, cigar = arrs
    ex.scan_batch(tid[:100], pos[:100], meta[:100], off[:101], cigar[:off[100]])
    o2 = (off[100:] - off[100]).astype(np.uint32)
    ex.scan_batch(tid[100:], pos[100:], meta[100:], o2, cigar[off[100]:], first_ordinal=100)
    g_tab = ex.junction_table()
    st = ex.stats()
    ex.close()
    o_tab, _ = run_oracle_batch(arrs, 0, contigs=("1",))
    tables_equal(g_tab, o_tab)
    assert st["table_grows"] >= 1


def test_add_junction_gtest_vectors():
    """tests/lib/junctions/test_junctions_extractor.cc:75-141 (JunctionName, AddJunction)."""
    rt = _rt()
    jc = rt.JunctionsExtractor(strandness=0)
    assert jc.get_new_junction_name() == "JUNC00000001"
    jc.add_junction(rt.Junction("chr1", 10000, 10200, 9500, 10700, "+"))
    assert jc.get_new_junction_name() == "JUNC00000002"
    jc.close()
    jc = rt.JunctionsExtractor(strandness=0)
    for args in [("chr1", 10000, 10200, 9900, 10300, "+"), ("chr1", 10000, 10200, 9500, 10200, "+"),
                 ("chr1", 10000, 10200, 9950, 10700, "+"), ("chr1", 8000, 8500, 7000, 10000, "+"),
                 ("chr1", 8000, 8500, 7000, 10000, "-")]:
        jc.add_junction(rt.Junction(*args))
    buf = io.StringIO()
    jc.print_all_junctions(buf)
    jc.close()
    expected = ("chr1\t7000\t10000\tJUNC00000002\t1\t+\t7000\t10000\t255,0,0\t2\t1000,1500\t0,1500\n"
                "chr1\t7000\t10000\tJUNC00000003\t1\t-\t7000\t10000\t255,0,0\t2\t1000,1500\t0,1500\n"
                "chr1\t9500\t10700\tJUNC00000001\t3\t+\t9500\t10700\t255,0,0\t2\t500,500\t0,700\n")
    assert buf.getvalue() == expected


GOLDENS = [
    (["-s", "XS"], "expected-a.out"), (["-s", "XS", "-a", "30"], "expected-a30.out"),
    (["-s", "RF"], "expected-stranded-a.out"), (["-s", "RF", "-a", "30"], "expected-stranded-a30.out"),
    (["-s", "XS", "-m", "8039", "-M", "8039"], "expected-i8039-I8039.out"),
    (["-s", "XS", "-r", "1:22405013-22405020"], "expected-r1:22405013-22405020.out"),
]


@pytest.mark.parametrize("args,golden", GOLDENS)
def test_reference_goldens(args, golden, hcc_bam, golden_dir, tmp_path):
    """tests/integration-test/test_junctions_extract.py:33-85 against the reference's own goldens."""
    rt = _rt()
    out = tmp_path / "extract.out"
    rc = rt.junctions_extract(["extract"] + args + ["-o", str(out), hcc_bam])
    assert rc == 0
    assert out.read_text() == open(os.path.join(golden_dir, "hcc1395", golden)).read()


def test_whole_bam_matches_oracle_fr_mode(hcc_bam):
    rt = _rt()
    ex = rt.JunctionsExtractor(hcc_bam, ".", 2, "XS", 8, 70, 500000)
    ex.identify_junctions_from_BAM()
    g_tab = ex.junction_table()
    st = ex.stats()
    ex.close()
    o = Oracle(8, 70, 500000, 2)
    o.extract_bam(hcc_bam)
    tables_equal(g_tab, o.table())
    assert st["reads"] == o.reads_seen() == 31678


def test_exit_codes(hcc_bam, tmp_path):
    """tests/integration-test/test_junctions_extract.py:87-109."""
    rt = _rt()
    out = str(tmp_path / "o")
    assert rt.junctions_extract(["extract", "-s", "XS", "-o", out]) == 1
    assert rt.junctions_extract(["extract", "-s", "XS", "-o", out, "does_not_exist.bam"]) == 1
    assert rt.junctions_extract(["extract", "-o", out]) == 1
    assert rt.junctions_extract(["extract", "-h"]) == 0


def _manifest(golden_dir):
    rows = []
    for line in open(os.path.join(golden_dir, "kat", "MANIFEST.tsv")):
        bam, out, args = line.rstrip("\n").split("\t")
        rows.append((bam, out, args.split()))
    return rows


def test_kat_goldens_from_the_reference(golden_dir, tmp_path):
    """Outputs of the unmodified reference on tests/golden/kat (make_golden.py): every op code, odd XS tags,
    multi-contig naming, records straddling tiny BGZF blocks, -r / -t / -a / -m / -M."""
    rt = _rt()
    n = 0
    for bam, out, args in _manifest(golden_dir):
        if args[0] == "ctor":
            continue
        dst = tmp_path / out
        rc = rt.junctions_extract(["extract"] + args + ["-o", str(dst), os.path.join(golden_dir, "kat", bam)])
        assert rc == 0
        assert dst.read_text() == open(os.path.join(golden_dir, "kat", out)).read(), (bam, args)
        n += 1
    assert n >= 15


def test_ctor_goldens_second_caller(golden_dir):
    """cis_splice_effects_identifier.cc:288-290: 8-arg ctor (min_intron := min_anchor) + unfiltered get_all_junctions."""
    rt = _rt()
    for bam, out, args in _manifest(golden_dir):
        if args[0] != "ctor":
            continue
        _, region, strandness, tag, anchor, min_intron, max_intron = args
        ex = rt.JunctionsExtractor.from_region(os.path.join(golden_dir, "kat", bam), region, int(strandness), tag,
                                               int(anchor), int(min_intron), int(max_intron))
        ex.identify_junctions_from_BAM()
        lines = "".join(f"{j.chrom}\t{j.thick_start}\t{j.thick_end}\t{j.name}\t{j.read_count}\t{j.strand}\t{j.start}\t{j.end}"
                        f"\t{int(j.has_left_min_anchor)}\t{int(j.has_right_min_anchor)}\n" for j in ex.get_all_junctions())
        ex.close()
        assert lines == open(os.path.join(golden_dir, "kat", out)).read(), args


def test_cis_splice_effects_junction_goldens(golden_dir):
    """The junction the reference's `cis-splice-effects identify -j` goldens hold for test_hcc1395.2.bam."""
    rt = _rt()
    bam = os.path.join(golden_dir, "hcc1395", "test_hcc1395.2.bam")
    for strandness, name in ((0, "expected-cis-splice-effects-identify-default-junctions.out"),
                             (1, "expected-cis-splice-effects-identify-default-stranded-junctions.out")):
        want = open(os.path.join(golden_dir, "hcc1395", name)).read().split("\t")
        ex = rt.JunctionsExtractor.from_region(bam, ".", strandness, "XS", 8, 70, 500000)
        ex.identify_junctions_from_BAM()
        js = ex.get_all_junctions()
        ex.close()
        hit = [j for j in js if str(j.thick_start) == want[1] and str(j.thick_end) == want[2]]
        assert hit and str(hit[0].read_count) == want[4] and hit[0].bed12().split("\t")[10:] == want[10:]


def test_cli_binary(golden_dir, hcc_bam, tmp_path):
    """The `regtools` executable (C++ shim over the C ABI): same goldens, same exit codes."""
    import subprocess
    exe = os.path.join(os.path.dirname(golden_dir), "..", "regtools_b200", "regtools")
    out = tmp_path / "cli.bed"
    p = subprocess.run([exe, "junctions", "extract", "-s", "XS", "-o", str(out), hcc_bam], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "Program:\tregtools" in p.stderr and "Minimum junction anchor length: 8" in p.stderr
    assert out.read_text() == open(os.path.join(golden_dir, "hcc1395", "expected-a.out")).read()
    p = subprocess.run([exe, "junctions", "extract", "-s", "RF", "-a", "30", hcc_bam], capture_output=True, text=True)
    assert p.returncode == 0 and p.stdout == open(os.path.join(golden_dir, "hcc1395", "expected-stranded-a30.out")).read()
    assert subprocess.run([exe, "junctions", "extract", "-s", "XS", "nope.bam"], capture_output=True).returncode == 1
    assert subprocess.run([exe, "junctions", "extract", hcc_bam], capture_output=True).returncode == 1
    assert subprocess.run([exe, "junctions", "extract", "-h"], capture_output=True).returncode == 0


def test_contig_sharded_handles_merge_to_whole_file(golden_dir):
    """Two shard handles (what two ranks run) + import on a host-only handle == single whole-file run."""
    rt = _rt()
    bam = os.path.join(golden_dir, "kat", "synth.bam")
    whole = rt.JunctionsExtractor(bam, ".", 0)
    whole.identify_junctions_from_BAM()
    want = io.StringIO(); whole.print_all_junctions(want)
    names = whole.contig_names()
    whole.close()
    merged = rt.JunctionsExtractor(bam, device=-1)
    merged.set_contigs(names)
    total = 0
    for r in (1, 0):
        ex = rt.JunctionsExtractor(bam, ".", 0, shard_rank=r, shard_world=2)
        ex.identify_junctions_from_BAM()
        t = ex.junction_table()
        total += ex.stats()["reads"]
        ex.close()
        merged.import_table(t)
    got = io.StringIO(); merged.print_all_junctions(got)
    assert got.getvalue() == want.getvalue()
    assert total == 6000


def test_large_generated_bam_property_checks(tmp_path):
    """Size-independent properties at a larger size: counts add up, output sorted, names a permutation,
    idempotent re-run, and equality with the oracle."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    bam = str(tmp_path / "big.bam")
    subprocess.check_call([os.path.join(root, "tools", "bamgen"), "gen", "--out", bam, "--config", "tiny", "--reads", "2000000",
                           "--seed", "77"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    rt = _rt()
    ex = rt.JunctionsExtractor(bam, ".", 0)
    ex.identify_junctions_from_BAM()
    t1 = ex.junction_table()
    st = ex.stats()
    ex.clear()
    ex.identify_junctions_from_BAM()
    t2 = ex.junction_table()
    ex.close()
    assert np.array_equal(t1, t2)
    assert st["reads"] == 2000000
    assert sorted(t1["name_index"].tolist()) == list(range(1, len(t1) + 1))
    key = list(zip(t1["tid"].tolist(), t1["thick_start"].tolist(), t1["thick_end"].tolist(), t1["name_index"].tolist()))
    order = {0: 0, 1: 1, 2: 2}                       # contigs "1" < "10" < "2" are already in tid order
    assert key == sorted(key, key=lambda k: (order[k[0]], k[1], k[2], k[3]))
    o = Oracle(8, 70, 500000, 0)
    o.extract_bam(bam)
    tables_equal(t1, o.table())
    assert int(t1["read_count"].sum()) == int(o.table()["read_count"].sum())


def test_full_size_c2_three_paths_agree():
    """BASELINE configs[1] at full size (10M reads, the bench's BAM): size-independent properties.  The device feeder
    (GPU inflate + record split), the host feeder and the kernel-level batch path must give the same table; names are a
    permutation of 1..U; order is compare_junctions'; read counts add up to the QC-passing candidates; the warp-pipelined
    scan kernel (variant 8) agrees; a sub-region agrees with the oracle run on that region."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    rt = _rt()
    bam = bench.ensure_bam("c2", 10_000_000, 6)
    dev = rt.JunctionsExtractor(bam, ".", 0, inflate_mode=2)
    dev.identify_junctions_from_BAM()
    t_dev, st_dev = dev.junction_table(), dev.stats()
    dev.close()
    host = rt.JunctionsExtractor(bam, ".", 0, inflate_mode=1)
    host.identify_junctions_from_BAM()
    t_host, st_host = host.junction_table(), host.stats()
    host.close()
    assert st_dev["reads"] == st_host["reads"] == 10_000_000 and st_dev["inflated_bytes"] == st_host["inflated_bytes"]
    assert st_dev["candidates"] == st_host["candidates"]
    assert np.array_equal(t_dev, t_host)
    ld = rt.JunctionsExtractor(bam, ".", 0, device=-1)
    arrs = ld.load_batch()
    ld.close()
    n_n = synth.count_n_ops(arrs[4])
    assert n_n == st_dev["candidates"]
    d = [torch.from_numpy(x.view(np.int32)).cuda() for x in arrs]
    for variant in (5, 8):
        ex = rt.JunctionsExtractor(bam, ".", 0, scan_variant=variant)
        ex.set_contigs(["chr1"])
        ex.scan_batch(*d, n_junction_ops=n_n)
        t = ex.junction_table()
        ex.close()
        assert np.array_equal(t, t_dev), variant
    U = len(t_dev)
    assert sorted(t_dev["name_index"].tolist()) == list(range(1, U + 1))
    key = list(zip(t_dev["thick_start"].tolist(), t_dev["thick_end"].tolist(), t_dev["name_index"].tolist()))
    assert key == sorted(key)
    # every N op whose intron length passes junction_qc is counted exactly once
    cig = arrs[4]
    ln = (cig[(cig & 0xF) == 3] >> 4).astype(np.int64)
    multi = np.diff(arrs[3].astype(np.int64)) > 1                  # all reads of this BAM are mapped to chr1
    assert multi.any()
    assert int(t_dev["read_count"].sum()) == int(np.count_nonzero((ln >= 70) & (ln <= 500000)))
    o = Oracle(8, 70, 500000, 0)
    o.extract_bam(bam, "chr1:100000000-100400000")
    sub = rt.JunctionsExtractor(bam, "chr1:100000000-100400000", 0)
    sub.identify_junctions_from_BAM()
    tables_equal(sub.junction_table(), o.table())
    sub.close()


@pytest.mark.parametrize("bam,csi", [("kat.bam", "kat.min14.csi"), ("kat.bam", "kat.min12.csi"), ("synth.bam", "synth.min14.csi")])
def test_csi_indexed_bam_reproduces_the_reference_goldens(bam, csi, golden_dir, tmp_path):
    """The same BAMs with only a .csi next to them (built by the reference's htslib): every golden of the manifest, through
    the host feeder and — whole-file runs — through the device feeder, whose record-walk seeds then come from the bin
    chunks alone (a CSI has no linear index)."""
    import shutil
    rt = _rt()
    src = str(tmp_path / bam)
    shutil.copy(os.path.join(golden_dir, "kat", bam), src)
    shutil.copy(os.path.join(golden_dir, "csi", csi), src + ".csi")
    n = 0
    for line in open(os.path.join(golden_dir, "kat", "MANIFEST.tsv")):
        b, out, args = line.rstrip("\n").split("\t")
        if b != bam or args.startswith("ctor"):
            continue
        want = open(os.path.join(golden_dir, "kat", out)).read()
        for mode in ((1, 2) if "-r" not in args.split() else (0,)):
            ex = rt.JunctionsExtractor(inflate_mode=mode)
            ex.parse_options(["extract"] + args.split() + [src])
            ex.identify_junctions_from_BAM()
            buf = io.StringIO()
            ex.print_all_junctions(buf)
            st = ex.stats()
            ex.close()
            assert buf.getvalue() == want, (csi, args, mode)
            if mode == 2 and bam == "synth.bam":
                assert st["host_parse_s"] == 0.0 and st["inflated_bytes"] > 0      # really the device feeder
            n += 1
    assert n >= 6
