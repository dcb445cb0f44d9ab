"""ctypes wrapper around oracle/_ref/libjx_oracle.so — TEST INFRASTRUCTURE ONLY.

The oracle is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
CPU-baseline legs import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "_ref", "libjx_oracle.so")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "regtools_ref")


def build():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
    return LIB


class OJunction(C.Structure):
    _fields_ = [("tid", C.c_int32), ("start", C.c_uint32), ("end", C.c_uint32), ("thick_start", C.c_uint32),
                ("thick_end", C.c_uint32), ("read_count", C.c_uint32), ("name_index", C.c_uint32),
                ("strand", C.c_uint8), ("left_ok", C.c_uint8), ("right_ok", C.c_uint8), ("pad", C.c_uint8)]


OJ_DTYPE = np.dtype([("tid", "<i4"), ("start", "<u4"), ("end", "<u4"), ("thick_start", "<u4"), ("thick_end", "<u4"),
                     ("read_count", "<u4"), ("name_index", "<u4"), ("strand", "u1"), ("left_ok", "u1"),
                     ("right_ok", "u1"), ("pad", "u1")])
OC_DTYPE = np.dtype([("start", "<u4"), ("end", "<u4"), ("thick_start", "<u4"), ("thick_end", "<u4"),
                     ("read_index", "<u8"), ("tid", "<i4"), ("k", "<u2"), ("strand", "u1"), ("pad", "u1")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        l = C.CDLL(LIB)
        l.jxo_new.restype = C.c_void_p
        l.jxo_new.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_char_p]
        l.jxo_free.argtypes = [C.c_void_p]
        l.jxo_set_contigs.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_char_p)]
        l.jxo_set_fasta.restype = C.c_int
        l.jxo_set_fasta.argtypes = [C.c_void_p, C.c_char_p]
        l.jxo_error.restype = C.c_char_p
        l.jxo_error.argtypes = [C.c_void_p]
        l.jxo_batch.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 5
        l.jxo_add.argtypes = [C.c_void_p, C.c_int32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint8]
        l.jxo_record_candidates.argtypes = [C.c_void_p, C.c_int]
        l.jxo_candidates.restype = C.c_size_t
        l.jxo_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        l.jxo_extract_bam.restype = C.c_int
        l.jxo_extract_bam.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_char_p)]
        l.jxo_reads_seen.restype = C.c_uint64
        l.jxo_reads_seen.argtypes = [C.c_void_p]
        l.jxo_count.restype = C.c_size_t
        l.jxo_count.argtypes = [C.c_void_p]
        l.jxo_get.restype = C.c_size_t
        l.jxo_get.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        l.jxo_write_bed12_path.argtypes = [C.c_void_p, C.c_char_p]
        l.jxo_enable_barcodes.argtypes = [C.c_void_p, C.c_char_p]
        l.jxo_set_read_barcode.argtypes = [C.c_void_p, C.c_char_p]
        l.jxo_batch_barcodes.argtypes = [C.c_void_p, C.c_uint32] + [C.c_void_p] * 6 + [C.POINTER(C.c_char_p)]
        l.jxo_barcodes_missing.restype = C.c_uint64
        l.jxo_barcodes_missing.argtypes = [C.c_void_p]
        l.jxo_write_barcode_replay_path.argtypes = [C.c_void_p, C.c_char_p]
        l.jxo_contig.restype = C.c_char_p
        l.jxo_contig.argtypes = [C.c_void_p, C.c_int32]
        _lib = l
    return _lib


class Oracle:
    """CPU restatement of JunctionsExtractor (oracle/jx_oracle.c)."""

    def __init__(self, min_anchor=8, min_intron=70, max_intron=500000, strandness=0, tag="XS", contigs=None, fasta=None,
                 barcodes=False):
        self.l = lib()
        self.h = C.c_void_p(self.l.jxo_new(min_anchor, min_intron, max_intron, strandness, tag.encode()))
        if barcodes:
            self.l.jxo_enable_barcodes(self.h, None)
        if contigs is not None:
            self.set_contigs(contigs)
        if fasta is not None and self.l.jxo_set_fasta(self.h, os.fsencode(fasta)):
            raise RuntimeError(f"cannot read {fasta}")

    def error(self):
        e = self.l.jxo_error(self.h)
        return e.decode() if e else None

    def set_contigs(self, names):
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self.l.jxo_set_contigs(self.h, len(names), arr)

    def close(self):
        if self.h:
            self.l.jxo_free(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def batch(self, tid, pos, meta, cig_off, cigar):
        tid, pos, meta, cig_off, cigar = (np.ascontiguousarray(x) for x in (tid, pos, meta, cig_off, cigar))
        self.l.jxo_batch(self.h, len(tid), tid.ctypes.data, pos.ctypes.data, meta.ctypes.data, cig_off.ctypes.data,
                         cigar.ctypes.data)

    def add(self, tid, start, end, ts, te, strand):
        self.l.jxo_add(self.h, tid, start & 0xFFFFFFFF, end & 0xFFFFFFFF, ts & 0xFFFFFFFF, te & 0xFFFFFFFF,
                       ord(strand) if isinstance(strand, str) else strand)

    def record_candidates(self, on=True):
        self.l.jxo_record_candidates(self.h, int(on))

    def candidates(self):
        n = self.l.jxo_candidates(self.h, None, 0)
        arr = np.zeros(n, OC_DTYPE)
        if n:
            self.l.jxo_candidates(self.h, arr.ctypes.data, n)
        return arr

    def extract_bam(self, bam, region="."):
        err = C.c_char_p()
        rc = self.l.jxo_extract_bam(self.h, os.fsencode(bam), region.encode(), C.byref(err))
        if rc:
            raise RuntimeError((err.value or b"error").decode())
        return rc

    def reads_seen(self):
        return self.l.jxo_reads_seen(self.h)

    def table(self):
        n = self.l.jxo_count(self.h)
        arr = np.zeros(n, OJ_DTYPE)
        if n:
            self.l.jxo_get(self.h, arr.ctypes.data, n)
        return arr

    def bed12(self):
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".bed") as f:
            self.l.jxo_write_bed12_path(self.h, f.name.encode())
            return open(f.name).read()


    # -b single-cell mode
    def set_read_barcode(self, bc):
        self.l.jxo_set_read_barcode(self.h, bc.encode())

    def batch_barcodes(self, tid, pos, meta, cig_off, cigar, bc, names):
        tid, pos, meta, cig_off, cigar, bc = (np.ascontiguousarray(x) for x in (tid, pos, meta, cig_off, cigar, bc))
        arr = (C.c_char_p * len(names))(*[n.encode() for n in names])
        self.l.jxo_batch_barcodes(self.h, len(tid), tid.ctypes.data, pos.ctypes.data, meta.ctypes.data, cig_off.ctypes.data,
                                  cigar.ctypes.data, bc.ctypes.data, arr)

    def barcodes_missing(self):
        return self.l.jxo_barcodes_missing(self.h)

    def barcodes(self):
        """Bytes of the reference's -b file: first-seen lists of the printed junctions (C oracle) replayed through a real
        std::unordered_map (oracle/bc_replay.cc)."""
        import tempfile
        exe = os.path.join(ORACLE_DIR, "_ref", "bc_replay")
        if not os.path.exists(exe):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "-s"])
        with tempfile.NamedTemporaryFile(suffix=".replay") as f:
            self.l.jxo_write_barcode_replay_path(self.h, f.name.encode())
            return subprocess.run([exe], stdin=open(f.name), capture_output=True, text=True, check=True).stdout


def ref_available():
    return os.path.exists(REF_BIN)


def ref_extract(bam, args):
    """Runs the UNMODIFIED reference (oracle/_ref/regtools_ref) and returns (rc, stdout)."""
    p = subprocess.run([REF_BIN, "junctions", "extract"] + list(args) + [bam], capture_output=True, text=True)
    return p.returncode, p.stdout
