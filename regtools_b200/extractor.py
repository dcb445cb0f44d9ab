"""Host-side mirror of the reference's JunctionsExtractor interface for the `junctions extract` path.

Same method names, argument meaning and error behaviour as
/root/reference/src/junctions/junctions_extractor.h:149-248 (class JunctionsExtractor) and
struct Junction (:39-112); everything forwards to the C ABI in libregtools_jx.so, which runs the
CUDA kernels.  PyTorch is used by callers only for device memory / streams / torch.distributed.
"""
import ctypes as C
import getopt
import os
import sys
from dataclasses import dataclass
from typing import Tuple, List, Optional, Sequence

import numpy as np

from . import _lib as L


class CmdlineHelpException(RuntimeError):
    """common::cmdline_help_exception (src/utils/common.h:96-101): `-h`, exit code 0."""


_MESSAGES = {
    L.RTJX_E_OPEN_BAM: "Unable to open BAM/SAM file.\n\n",
    L.RTJX_E_OPEN_INDEX: "Unable to open BAM/SAM index. Make sure alignments are indexed\n\n",
    L.RTJX_E_REGION: "Unable to iterate to region within BAM.\n\n",
}

USAGE = (
    "Usage:\t\tregtools junctions extract [options] indexed_alignments.bam\n"
    "Options:\n"
    "\t\t-a INT\tMinimum anchor length. Junctions which satisfy a minimum \n"
    "\t\t\t anchor length on both sides are reported. [8]\n"
    "\t\t-m INT\tMinimum intron length. [70]\n"
    "\t\t-M INT\tMaximum intron length. [500000]\n"
    "\t\t-o FILE\tThe file to write output to. [STDOUT]\n"
    "\t\t-r STR\tThe region to identify junctions \n"
    "\t\t\t in \"chr:start-end\" format. Entire BAM by default.\n"
    "\t\t-s INT\tStrandness mode \n"
    "\t\t\t XS, use XS tags provided by aligner; RF, first-strand; FR, second-strand. REQUIRED\n"
    "\t\t-t STR\tTag used in bam to label strand. [XS]\n"
    "\t\t-b STR\tThe file containing the barcodes of interest for single cell data.\n"
    "\n")


@dataclass
class Junction:
    """struct Junction (junctions_extractor.h:39-112) without the barcode map."""
    chrom: str = ""
    start: int = 0
    end: int = 0
    thick_start: int = 0
    thick_end: int = 0
    strand: str = "?"
    name: str = "NA"
    read_count: int = 0
    has_left_min_anchor: bool = False
    has_right_min_anchor: bool = False
    color: str = "255,0,0"
    nblocks: int = 2

    def bed12(self) -> str:
        """Junction::print (junctions_extractor.h:90-98); uint32 arithmetic."""
        m = 0xFFFFFFFF
        return (f"{self.chrom}\t{self.thick_start}\t{self.thick_end}\t{self.name}\t{self.read_count}\t{self.strand}"
                f"\t{self.thick_start}\t{self.thick_end}\t{self.color}\t{self.nblocks}"
                f"\t{(self.start - self.thick_start) & m},{(self.thick_end - self.end) & m}"
                f"\t0,{(self.end - self.thick_start) & m}\n")


class JunctionsExtractor:
    """B200-backed drop-in for the reference class of the same name."""

    def __init__(self, bam: str = "NA", region: str = ".", strandness: int = -1, strand_tag: str = "XS",
                 min_anchor_length: int = 8, min_intron_length: int = 70, max_intron_length: int = 500000,
                 ref: str = "NA", *, device: int = 0, n_threads: int = 0, batch_reads: int = 0,
                 shard_rank: int = 0, shard_world: int = 1, profile: bool = False, table_log2: int = 0,
                 inflate_mode: int = 0, scan_variant: int = 0, scan_cfg: int = 0,
                 _ctor8: bool = False):
        self.bam_ = bam
        self.region_ = region
        self.strandness_ = strandness
        self.strand_tag_ = strand_tag
        self.min_anchor_length_ = min_anchor_length & 0xFFFFFFFF
        # 8-arg ctor quirk (junctions_extractor.h:199-200): min_intron_length_ is initialised from
        # min_anchor_length1; reproduced by from_region().
        self.min_intron_length_ = (min_anchor_length if _ctor8 else min_intron_length) & 0xFFFFFFFF
        self.max_intron_length_ = max_intron_length & 0xFFFFFFFF
        self.ref_ = ref
        self.output_file_ = "NA"
        self.output_barcodes_file_ = "NA"
        self.barcode_tag_ = "CB"                    # junctions_extractor.h:192,204
        self._opts = dict(device=device, n_threads=n_threads, batch_reads=batch_reads, shard_rank=shard_rank,
                          shard_world=shard_world, profile=int(profile), table_log2=table_log2,
                          inflate_mode=inflate_mode, scan_variant=scan_variant, scan_cfg=scan_cfg)
        self._h = None

    @classmethod
    def from_region(cls, bam1, region1, strandness1, strand_tag1, min_anchor_length1, min_intron_length1,
                    max_intron_length1, ref1="NA", **kw):
        """The 8-argument constructor used by cis-splice-effects
        (cis_splice_effects_identifier.cc:288), including its min_intron := min_anchor quirk."""
        return cls(bam1, region1, strandness1, strand_tag1, min_anchor_length1, min_intron_length1,
                   max_intron_length1, ref1, _ctor8=True, **kw)

    # ------------------------------------------------------------------ handle plumbing
    def _handle(self):
        if self._h is None:
            p = L.Params()
            L.lib.rtjx_params_default(C.byref(p))
            p.bam = None if self.bam_ in ("NA", "") else os.fsencode(self.bam_)
            p.region = self.region_.encode()
            p.strand_tag = self.strand_tag_.encode()
            p.fasta = None if self.ref_ == "NA" else os.fsencode(self.ref_)
            p.barcode_out = None if self.output_barcodes_file_ == "NA" else os.fsencode(self.output_barcodes_file_)
            p.strandness = max(self.strandness_, 0)
            p.min_anchor, p.min_intron, p.max_intron = (self.min_anchor_length_, self.min_intron_length_,
                                                        self.max_intron_length_)
            for k, v in self._opts.items():
                setattr(p, k, v)
            self._params = p  # keeps the byte strings alive
            h = C.c_void_p()
            rc = L.lib.rtjx_create(C.byref(p), C.byref(h))
            if rc != L.RTJX_OK:
                raise RuntimeError(f"rtjx_create: {L.lib.rtjx_last_error(None).decode()} ({rc})")
            self._h = h
        return self._h

    def _check(self, rc):
        if rc >= 0:
            return rc
        msg = _MESSAGES.get(rc) or (L.lib.rtjx_last_error(self._h) or b"").decode() or L.lib.rtjx_strerror(rc).decode()
        if rc == L.RTJX_E_OPEN_BAM:                  # htslib's own stderr line before the reference's exception (hts_open_format)
            sys.stderr.write(f"[E::hts_open_format] fail to open file '{self.bam_}'\n")
        raise RuntimeError(msg)

    def close(self):
        if self._h is not None:
            L.lib.rtjx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ reference interface
    def usage(self, out=sys.stderr):
        out.write(USAGE)
        return 0

    def parse_options(self, argv: Sequence[str]) -> int:
        """parse_options (junctions_extractor.cc:42-122). argv[0] is the sub-command name."""
        try:
            opts, rest = getopt.getopt(list(argv[1:]), "ha:m:M:o:r:t:s:b:")
        except getopt.GetoptError as e:
            _glibc_getopt_line(argv[0], e)
            self.usage()
            raise RuntimeError("Error parsing inputs!(1)\n\n")
        atoi = _atoi
        for o, a in opts:
            if o == "-h":
                raise CmdlineHelpException(USAGE)
            elif o == "-a":
                self.min_anchor_length_ = atoi(a) & 0xFFFFFFFF
            elif o == "-m":
                self.min_intron_length_ = atoi(a) & 0xFFFFFFFF
            elif o == "-M":
                self.max_intron_length_ = atoi(a) & 0xFFFFFFFF
            elif o == "-o":
                self.output_file_ = a
            elif o == "-r":
                self.region_ = a
            elif o == "-t":
                self.strand_tag_ = a
            elif o == "-s":
                modes = {"XS": 0, "RF": 1, "FR": 2, "intron-motif": 3}
                if a not in modes:
                    raise RuntimeError("Unrecognized strandness argument!\n\n")
                self.strandness_ = modes[a]
            elif o == "-b":
                self.output_barcodes_file_ = a
        if len(rest) >= 1:
            self.bam_ = rest.pop(0)
        if len(rest) >= 1:
            self.ref_ = rest.pop(0)
        if rest or self.bam_ == "NA":
            self.usage()
            raise RuntimeError("Error parsing inputs!(2)\n\n")
        if self.strandness_ == -1:
            self.usage()
            raise RuntimeError("Please supply strandness mode with '-s' option!\n\n")
        if self.strandness_ == 3 and self.ref_ == "NA":
            self.usage()
            raise RuntimeError("Strandness mode 'intron-motif' requires a fasta file!\n\n")
        e = sys.stderr
        e.write(f"Minimum junction anchor length: {self.min_anchor_length_}\n")
        e.write(f"Minimum intron length: {self.min_intron_length_}\n")
        e.write(f"Maximum intron length: {self.max_intron_length_}\n")
        e.write(f"Alignment: {self.bam_}\n")
        e.write(f"Output file: {self.output_file_}\n")
        if self.output_barcodes_file_ != "NA":
            e.write(f"Barcode file: {self.output_barcodes_file_}\n")
        e.write("\n")
        return 0

    def get_bam(self) -> str:
        return self.bam_

    def identify_junctions_from_BAM(self) -> int:
        """identify_junctions_from_BAM (junctions_extractor.cc:500-535)."""
        if not self.bam_:
            return 0
        self._check(L.lib.rtjx_run(self._handle()))
        if self.output_barcodes_file_ != "NA":
            # set_junction_barcode (junctions_extractor.cc:369-372); aln->id is never set, the reference prints 0
            _, n_missing = self.barcode_stats()
            for _ in range(n_missing):
                sys.stderr.write(f"WARNING: No {self.barcode_tag_} tag found for alignment (id = 0)\n")
        return 0

    # ------------------------------------------------------------------ -b single-cell barcodes
    def barcode_stats(self):
        """(distinct barcode values seen incl. '?', n_cigar > 1 alignments without the tag)."""
        nb, nm = C.c_uint64(), C.c_uint64()
        self._check(L.lib.rtjx_barcode_stats(self._handle(), C.byref(nb), C.byref(nm)))
        return int(nb.value), int(nm.value)

    def barcode_names(self) -> List[str]:
        h = self._handle()
        return [L.lib.rtjx_barcode_name(h, i).decode() for i in range(self.barcode_stats()[0])]

    def load_barcodes(self) -> np.ndarray:
        """Host feeder only: the barcode dictionary id of every alignment the iterator visits (0 for n_cigar <= 1)."""
        h = self._handle()
        n = self._check(L.lib.rtjx_load_barcodes(h, None, 0))
        ids = np.zeros(n, dtype=np.uint32)
        if n:
            self._check(L.lib.rtjx_load_barcodes(h, ids.ctypes.data, n))
        return ids

    def print_barcodes(self, out=None) -> None:
        """Junction::print_barcodes (junctions_extractor.h:99-111) for every printed junction, in print order: to the
        -b file (junctions_extractor.cc:255-257,272-273), or to `out` when given."""
        h = self._handle()
        if out is None:
            fd = os.open(self.output_barcodes_file_, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            try:
                self._check(L.lib.rtjx_write_barcodes(h, fd))
            finally:
                os.close(fd)
            return
        out.write(self._capture(lambda w: L.lib.rtjx_write_barcodes(h, w)))

    def _capture(self, writer) -> str:
        r, w = os.pipe()
        import threading
        chunks = []
        t = threading.Thread(target=lambda: chunks.append(_read_all(r)))
        t.start()
        try:
            self._check(writer(w))
        finally:
            os.close(w)
            t.join()
            os.close(r)
        return b"".join(chunks).decode()

    def identify_junctions_in_regions(self, regions: Sequence[str]) -> List[np.ndarray]:
        """Batched form of the per-variant loop of cis-splice-effects (cis_splice_effects_identifier.cc:267-311): the
        junction table of every region from one pass over the BAM.  Element i equals what
        ``JunctionsExtractor(bam, regions[i], ...)`` + ``identify_junctions_from_BAM`` + ``junction_table`` returns."""
        h = self._handle()
        arr = (C.c_char_p * max(len(regions), 1))(*[r.encode() for r in regions])
        self._check(L.lib.rtjx_run_regions(h, arr, len(regions)))
        out = []
        for i in range(len(regions)):
            n = self._check(L.lib.rtjx_region_count(h, i))
            t = np.zeros(n, dtype=JUNCTION_DTYPE)
            if n:
                self._check(L.lib.rtjx_region_get(h, i, t.ctypes.data_as(C.POINTER(L.Junction)), n))
            out.append(t)
        return out

    def unique_junctions_in_windows(self, regions: Sequence[str], windows: Sequence[Tuple[int, int]]):
        """The junction side of `cis-splice-effects identify`'s per-variant loop (cis_splice_effects_identifier.cc:288-299):
        ``regions[i]`` / ``windows[i] = (cis_effect_start, cis_effect_end)`` describe variant i.  Returns
        (table, first_region, variants): the reference's ``unique_junctions_`` set in its iteration order (ordered and deduplicated
        by contig name, start, end — strand-blind, first insert wins, junctions_annotator.h:155-177), the variant whose insert
        won, and for every unique junction the ascending list of variants whose window holds it (``junction_to_variant_``)."""
        self.identify_junctions_in_regions(regions)
        h = self._handle()
        n = len(regions)
        ws = (C.c_uint32 * max(n, 1))(*[int(w[0]) & 0xFFFFFFFF for w in windows])
        we = (C.c_uint32 * max(n, 1))(*[int(w[1]) & 0xFFFFFFFF for w in windows])
        self._check(L.lib.rtjx_unique_junctions(h, ws, we, n))
        u = self._check(L.lib.rtjx_unique_count(h))
        t = np.zeros(u, dtype=JUNCTION_DTYPE)
        first = np.zeros(u, dtype=np.uint32)
        if u:
            self._check(L.lib.rtjx_unique_get(h, t.ctypes.data_as(C.POINTER(L.Junction)), first.ctypes.data_as(C.POINTER(C.c_uint32)), u))
        variants = []
        for i in range(u):
            m = self._check(L.lib.rtjx_unique_regions(h, i, None, 0))
            v = np.zeros(m, dtype=np.uint32)
            self._check(L.lib.rtjx_unique_regions(h, i, v.ctypes.data_as(C.POINTER(C.c_uint32)), m))
            variants.append(v.tolist())
        return t, first, variants

    def get_new_junction_name(self) -> str:
        """get_new_junction_name (junctions_extractor.cc:152-157)."""
        n = self._check(L.lib.rtjx_count(self._handle()))
        return "JUNC%08d" % (n + 1)

    def add_junction(self, j: Junction) -> int:
        """add_junction (junctions_extractor.cc:174-235)."""
        h = self._handle()
        c = L.Candidate()
        c.tid = L.lib.rtjx_intern_contig(h, j.chrom.encode())
        c.start, c.end = j.start & 0xFFFFFFFF, j.end & 0xFFFFFFFF
        c.thick_start, c.thick_end = j.thick_start & 0xFFFFFFFF, j.thick_end & 0xFFFFFFFF
        c.strand = ord(j.strand[0]) if j.strand else ord("?")
        self._check(L.lib.rtjx_add(h, C.byref(c), 1))
        return 0

    def junction_table(self) -> np.ndarray:
        """All junctions (sorted, unfiltered) as a structured numpy array of rtjx_junction."""
        h = self._handle()
        n = self._check(L.lib.rtjx_count(h))
        arr = np.empty(n, dtype=JUNCTION_DTYPE)          # rtjx_get fills every byte (40-byte records, pad included)
        if n:
            self._check(L.lib.rtjx_get(h, arr.ctypes.data_as(C.POINTER(L.Junction)), n))
        return arr

    def get_all_junctions(self) -> List[Junction]:
        """get_all_junctions (junctions_extractor.cc:238-246): sorted, NOT anchor-filtered."""
        h = self._handle()
        out = []
        for r in self.junction_table():
            out.append(Junction(
                chrom=L.lib.rtjx_contig(h, int(r["tid"])).decode(), start=int(r["start"]), end=int(r["end"]),
                thick_start=int(r["thick_start"]), thick_end=int(r["thick_end"]), strand=chr(int(r["strand"])),
                name="JUNC%08d" % int(r["name_index"]), read_count=int(r["read_count"]),
                has_left_min_anchor=bool(r["left_ok"]), has_right_min_anchor=bool(r["right_ok"])))
        return out

    def print_all_junctions(self, out=None) -> None:
        """print_all_junctions (junctions_extractor.cc:249-280): -o file if set, else `out`/stdout."""
        h = self._handle()
        if self.output_barcodes_file_ != "NA":
            self.print_barcodes()
        if self.output_file_ != "NA":
            try:
                fd = os.open(self.output_file_, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644)
            except OSError:
                fd = -1        # an ofstream that failed to open: every junction goes to `out` (junctions_extractor.cc:269-272)
            if fd >= 0:
                try:
                    self._check(L.lib.rtjx_write_bed12(h, fd))
                finally:
                    os.close(fd)
                return
        if out is None:
            sys.stdout.flush()
            self._check(L.lib.rtjx_write_bed12(h, 1))
            return
        r, w = os.pipe()
        # small outputs only (tests); large ones should go through -o / a real fd
        import threading
        chunks = []
        t = threading.Thread(target=lambda: chunks.append(_read_all(r)))
        t.start()
        try:
            self._check(L.lib.rtjx_write_bed12(h, w))
        finally:
            os.close(w)
            t.join()
            os.close(r)
        out.write(b"".join(chunks).decode())

    # ------------------------------------------------------------------ batch level (kernels)
    def intern_barcode(self, barcode: str) -> int:
        """-b handles: dictionary id of a barcode string, the value for scan_batch's `bc` column."""
        return self._check(L.lib.rtjx_intern_barcode(self._handle(), barcode.encode()))

    def scan_batch(self, tid, pos, meta, cig_off, cigar, first_ordinal: int = 0, n_junction_ops: int = 0,
                   stream: Optional[int] = None, bc=None) -> int:
        """parse_alignment_into_junctions over a SoA batch.  Arrays are numpy (host) or torch CUDA
        tensors (device, used in place); dtypes int32/int32/uint32-compatible, see include/rtjx.h."""
        h = self._handle()
        b = L.Batch()
        is_dev = hasattr(tid, "is_cuda") and tid.is_cuda
        n_reads = int(tid.shape[0])
        n_ops = int(cigar.shape[0])
        if int(cig_off.shape[0]) != n_reads + 1:
            raise ValueError("cig_off must have n_reads + 1 entries")
        b.n_reads, b.n_ops, b.first_ordinal, b.n_junction_ops = n_reads, n_ops, first_ordinal, n_junction_ops
        ptr = (lambda t: t.data_ptr()) if is_dev else (lambda a: a.ctypes.data)
        if not is_dev:
            tid, pos, meta, cig_off, cigar = (np.ascontiguousarray(x) for x in (tid, pos, meta, cig_off, cigar))
            for x in (tid, pos, meta, cig_off, cigar):
                if x.dtype.itemsize != 4:
                    raise ValueError("batch arrays must be 32-bit")
        b.tid, b.pos, b.meta, b.cig_off, b.cigar = ptr(tid), ptr(pos), ptr(meta), ptr(cig_off), ptr(cigar) if n_ops else None
        if bc is not None:                               # -b handles: per-alignment barcode ids (same residency as the batch)
            if not is_dev:
                bc = np.ascontiguousarray(bc, dtype=np.uint32)
            b.bc = ptr(bc)
        self._keep = (tid, pos, meta, cig_off, cigar, bc)
        self._check(L.lib.rtjx_scan_batch(h, C.byref(b), L.RTJX_LOC_DEVICE if is_dev else L.RTJX_LOC_HOST,
                                          C.c_void_p(stream) if stream else None))
        return 0

    def finalize(self, stream: Optional[int] = None):
        self._check(L.lib.rtjx_finalize(self._handle(), C.c_void_p(stream) if stream else None))

    def clear(self):
        self._check(L.lib.rtjx_clear(self._handle()))

    def load_batch(self):
        """Host feeder only: the SoA arrays of every alignment the iterator visits."""
        h = self._handle()
        nr, no = C.c_uint64(), C.c_uint64()
        self._check(L.lib.rtjx_load_batch(h, C.byref(nr), C.byref(no), None, None, None, None, None))
        tid = np.empty(nr.value, np.int32); pos = np.empty(nr.value, np.int32)
        meta = np.empty(nr.value, np.uint32); off = np.empty(nr.value + 1, np.uint32)
        cig = np.empty(max(no.value, 1), np.uint32)
        self._check(L.lib.rtjx_load_batch(h, C.byref(nr), C.byref(no), tid.ctypes.data, pos.ctypes.data,
                                          meta.ctypes.data, off.ctypes.data, cig.ctypes.data))
        return tid, pos, meta, off, cig[:no.value]

    # ------------------------------------------------------------------ multi-GPU exchange (NCCL inside the library)
    @staticmethod
    def comm_unique_id() -> bytes:
        """rtjx_comm_unique_id: the 128 bytes rank 0 hands to every other rank before comm_init."""
        buf = C.create_string_buffer(128)
        rc = L.lib.rtjx_comm_unique_id(buf)
        if rc != L.RTJX_OK:
            raise RuntimeError(f"rtjx_comm_unique_id: {L.lib.rtjx_last_error(None).decode()} ({rc})")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int) -> None:
        """rtjx_comm_init: joins the process-wide NCCL communicator (once per process; later handles reuse it)."""
        self._check(L.lib.rtjx_comm_init(self._handle(), unique_id, rank, world))

    def gather(self, root: int = 0) -> None:
        """rtjx_gather: all shard tables to the root over NCCL; the root then holds the merged, ranked, sorted table."""
        self._check(L.lib.rtjx_gather(self._handle(), root))

    def stage_bam(self) -> None:
        """Copies the compressed BAM into device memory; later runs of this extractor read it from there (rtjx_stage_bam)."""
        self._check(L.lib.rtjx_stage_bam(self._handle()))

    def inflate_file(self, max_blocks: int = 0) -> bytes:
        """Device BGZF inflate of the BAM (test hook): returns the inflated byte stream."""
        h = self._handle()
        n = C.c_uint64()
        self._check(L.lib.rtjx_inflate_file(h, max_blocks, None, 0, C.byref(n)))
        buf = np.empty(max(n.value, 1), np.uint8)
        self._check(L.lib.rtjx_inflate_file(h, max_blocks, buf.ctypes.data, n.value, C.byref(n)))
        return buf[:n.value].tobytes()

    def import_table(self, table: np.ndarray):
        table = np.ascontiguousarray(table, dtype=JUNCTION_DTYPE)
        self._check(L.lib.rtjx_import(self._handle(), table.ctypes.data_as(C.POINTER(L.Junction)), len(table)))

    def contig_names(self) -> List[str]:
        h = self._handle()
        return [L.lib.rtjx_contig(h, i).decode() for i in range(L.lib.rtjx_n_contigs(h))]

    def set_contigs(self, names: Sequence[str]):
        h = self._handle()
        for n in names:
            L.lib.rtjx_intern_contig(h, n.encode())

    def stats(self) -> dict:
        s = L.Stats()
        L.lib.rtjx_get_stats(self._handle(), C.byref(s))
        return {f: getattr(s, f) for f, _ in L.Stats._fields_}

    def reset_stats(self):
        L.lib.rtjx_reset_stats(self._handle())


JUNCTION_DTYPE = np.dtype([
    ("tid", "<i4"), ("start", "<u4"), ("end", "<u4"), ("thick_start", "<u4"), ("thick_end", "<u4"),
    ("read_count", "<u4"), ("name_index", "<u4"), ("strand", "u1"), ("left_ok", "u1"), ("right_ok", "u1"),
    ("pad", "u1"), ("first_ord", "<u8")])
assert JUNCTION_DTYPE.itemsize == C.sizeof(L.Junction)


def _read_all(fd):
    out = []
    while True:
        b = os.read(fd, 1 << 20)
        if not b:
            return b"".join(out)
        out.append(b)


def _atoi(s: str) -> int:
    """C atoi: leading whitespace, optional sign, digits; garbage -> 0."""
    s = s.lstrip()
    i, sign = 0, 1
    if s[:1] in "+-" and s[:1]:
        sign = -1 if s[0] == "-" else 1
        i = 1
    j = i
    while j < len(s) and s[j].isdigit():
        j += 1
    return sign * int(s[i:j]) if j > i else 0


def plan_shards(bam: str, world: int) -> List[int]:
    """Contig -> shard assignment (rtjx_plan_shards)."""
    n = L.lib.rtjx_plan_shards(os.fsencode(bam), world, None, 0)
    if n < 0:
        raise RuntimeError(_MESSAGES.get(n, L.lib.rtjx_strerror(n).decode()))
    arr = (C.c_int32 * max(n, 1))()
    L.lib.rtjx_plan_shards(os.fsencode(bam), world, arr, n)
    return list(arr[:n])


ANNOTATE_USAGE = ("Usage:\t\tregtools junctions annotate [options] junctions.bed ref.fa annotations.gtf\n"
                  "Options:\t-S include single exon genes\n"
                  "\t\t-o FILE\tThe file to write output to. [STDOUT]\n\n")


class JunctionsAnnotator:
    """`regtools junctions annotate` over the C ABI (rtjx_annotate): the options and texts of the reference's
    JunctionsAnnotator (junctions_annotator.cc:405-456); the per-line loop of junctions_main.cc:68-82 runs batched on the
    device."""

    def __init__(self, junctions: str = "", ref: str = "NA", gtf: str = "", *, device: int = 0):
        self.junctions_, self.ref_, self.gtf_ = junctions, ref, gtf
        self.skip_single_exon_genes_ = True
        self.output_file_ = "NA"
        self.device_ = device

    def usage(self, out=None) -> int:
        (out or sys.stderr).write(ANNOTATE_USAGE)
        return 0

    def gtf_file(self) -> str:
        """gtf_file (junctions_annotator.cc:400-402)."""
        return self.gtf_

    def parse_options(self, argv: Sequence[str]) -> int:
        """argv[0] is the sub-command name, as in the reference's getopt call."""
        try:
            opts, args = getopt.getopt(list(argv[1:]), "So:h")
        except getopt.GetoptError as e:
            _glibc_getopt_line(argv[0], e)
            self.usage()
            raise RuntimeError("Error parsing inputs!(1)\n\n")
        for o, a in opts:
            if o == "-S":
                self.skip_single_exon_genes_ = False
            elif o == "-o":
                self.output_file_ = a
            elif o == "-h":
                raise CmdlineHelpException(ANNOTATE_USAGE)
        if len(args) >= 3:
            self.junctions_, self.ref_, self.gtf_ = args[0], args[1], args[2]
            args = args[3:]
        if args or self.ref_ == "NA" or not self.junctions_ or not self.gtf_:
            self.usage()
            raise RuntimeError("Error parsing inputs!(2)\n\n")
        e = sys.stderr
        e.write(f"Reference: {self.ref_}\nGTF: {self.gtf_}\nJunctions: {self.junctions_}\n")
        if self.skip_single_exon_genes_:
            e.write("Skipping single exon genes.\n")
        if self.output_file_ != "NA":
            e.write(f"Output file: {self.output_file_}\n")
        e.write("\n")
        return 0

    def annotate_all(self, out_fd: Optional[int] = None, chatter_fd: int = -1) -> int:
        """Header + one line per junction to -o / out_fd / stdout; returns the number of annotated lines."""
        p = L.AnnotateParams()
        L.lib.rtjx_annotate_params_default(C.byref(p))
        p.junctions_bed, p.fasta, p.gtf = os.fsencode(self.junctions_), os.fsencode(self.ref_), os.fsencode(self.gtf_)
        p.include_single_exon = 0 if self.skip_single_exon_genes_ else 1
        p.device = self.device_
        p.chatter_fd = chatter_fd
        fd = out_fd
        if fd is None:
            if self.output_file_ != "NA":
                fd, p.out_path = -1, os.fsencode(self.output_file_)
            else:
                sys.stdout.flush()
                fd = 1
        err = C.create_string_buffer(1024)
        n = C.c_uint64()
        rc = L.lib.rtjx_annotate(C.byref(p), fd, C.byref(n), err, len(err))
        if rc != L.RTJX_OK:
            raise RuntimeError(err.value.decode() or L.lib.rtjx_strerror(rc).decode())
        return int(n.value)


def _glibc_getopt_line(prog: str, e: "getopt.GetoptError") -> None:
    """The line glibc's getopt() itself writes to stderr before the reference's default: branch runs."""
    what = "option requires an argument" if "requires argument" in e.msg else "invalid option"
    sys.stderr.write(f"{prog}: {what} -- '{e.opt}'\n")


def junctions_annotate(argv: Sequence[str], device: int = 0) -> int:
    """junctions_annotate (src/junctions/junctions_main.cc:61-92): exit code 0 / 1."""
    anno = JunctionsAnnotator(device=device)
    try:
        anno.parse_options(argv)
        sys.stderr.flush()
        anno.annotate_all(chatter_fd=2)
    except CmdlineHelpException as e:
        sys.stderr.write(str(e) + "\n")
        return 0
    except RuntimeError as e:
        msg = str(e)
        # two GtfParser failures are `cerr << text; exit(1)` without a line end (gtf_parser.cc:52-55,202-206)
        bare = msg.startswith("\nUnable to open GTF file.") or msg.startswith("Undefined strand for exon")
        sys.stderr.write(msg if bare else msg + "\n")
        return 1
    return 0


def junctions_extract(argv: Sequence[str]) -> int:
    """junctions_extract (src/junctions/junctions_main.cc:45-59): exit code 0 / 1."""
    ex = JunctionsExtractor()
    try:
        ex.parse_options(argv)
        ex.identify_junctions_from_BAM()
        ex.print_all_junctions()
    except CmdlineHelpException as e:
        sys.stderr.write(str(e) + "\n")
        return 0
    except RuntimeError as e:
        sys.stderr.write(str(e) + "\n")
        return 1
    finally:
        ex.close()
    return 0
