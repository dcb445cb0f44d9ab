"""ctypes binding of libregtools_jx.so (C ABI declared in include/rtjx.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``make -C regtools_b200/csrc``.
There is no Python or CPU fallback: if the library is missing this module raises at import.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libregtools_jx.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C regtools_b200/csrc`. regtools_b200 has no fallback implementation.")

lib = C.CDLL(LIB_PATH)

RTJX_OK = 0
RTJX_E_ARG, RTJX_E_OPEN_BAM, RTJX_E_OPEN_INDEX, RTJX_E_REGION = -1, -2, -3, -4
RTJX_E_CUDA, RTJX_E_UNSUPPORTED, RTJX_E_NOMEM, RTJX_E_STATE, RTJX_E_IO = -5, -6, -7, -8, -9
RTJX_LOC_HOST, RTJX_LOC_DEVICE = 0, 1


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("bam", C.c_char_p), ("region", C.c_char_p), ("strand_tag", C.c_char_p),
        ("fasta", C.c_char_p), ("barcode_out", C.c_char_p),
        ("strandness", C.c_int32),
        ("min_anchor", C.c_uint32), ("min_intron", C.c_uint32), ("max_intron", C.c_uint32),
        ("device", C.c_int32), ("n_threads", C.c_int32),
        ("batch_reads", C.c_uint32), ("table_log2", C.c_uint32),
        ("shard_rank", C.c_int32), ("shard_world", C.c_int32),
        ("inflate_mode", C.c_int32), ("profile", C.c_int32),
        ("scan_variant", C.c_int32), ("scan_cfg", C.c_int32),
        ("barcode_tag", C.c_char_p),
    ]


class Junction(C.Structure):
    _fields_ = [
        ("tid", C.c_int32), ("start", C.c_uint32), ("end", C.c_uint32),
        ("thick_start", C.c_uint32), ("thick_end", C.c_uint32),
        ("read_count", C.c_uint32), ("name_index", C.c_uint32),
        ("strand", C.c_uint8), ("left_ok", C.c_uint8), ("right_ok", C.c_uint8), ("pad", C.c_uint8),
        ("first_ord", C.c_uint64),
    ]


class Candidate(C.Structure):
    _fields_ = [
        ("tid", C.c_int32), ("start", C.c_uint32), ("end", C.c_uint32),
        ("thick_start", C.c_uint32), ("thick_end", C.c_uint32),
        ("strand", C.c_uint8), ("pad", C.c_uint8 * 3),
    ]


class Batch(C.Structure):
    _fields_ = [
        ("n_reads", C.c_uint32), ("n_ops", C.c_uint32), ("first_ordinal", C.c_uint64),
        ("n_junction_ops", C.c_uint32), ("reserved", C.c_uint32),
        ("tid", C.c_void_p), ("pos", C.c_void_p), ("meta", C.c_void_p),
        ("cig_off", C.c_void_p), ("cigar", C.c_void_p), ("bc", C.c_void_p),
    ]


class AnnotateParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("junctions_bed", C.c_char_p), ("fasta", C.c_char_p), ("gtf", C.c_char_p),
        ("include_single_exon", C.c_int32), ("device", C.c_int32), ("chatter_fd", C.c_int32),
        ("out_path", C.c_char_p),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("reads", C.c_uint64), ("cigar_ops", C.c_uint64), ("candidates", C.c_uint64),
        ("batches", C.c_uint64), ("kernel_launches", C.c_uint64),
        ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("bgzf_blocks", C.c_uint64), ("compressed_bytes", C.c_uint64), ("inflated_bytes", C.c_uint64),
        ("scan_ms", C.c_double), ("merge_ms", C.c_double), ("finalize_ms", C.c_double),
        ("inflate_kernel_ms", C.c_double),
        ("host_inflate_s", C.c_double), ("host_parse_s", C.c_double), ("host_wait_s", C.c_double),
        ("total_s", C.c_double),
        ("table_slots", C.c_uint32), ("table_grows", C.c_uint32),
    ]


# every symbol include/rtjx.h declares, with its signature
SIGNATURES = {
    "rtjx_params_default": (None, [C.POINTER(Params)]),
    "rtjx_create": (C.c_int, [C.POINTER(Params), C.POINTER(C.c_void_p)]),
    "rtjx_destroy": (None, [C.c_void_p]),
    "rtjx_run": (C.c_int, [C.c_void_p]),
    "rtjx_run_regions": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.c_size_t]),
    "rtjx_region_count": (C.c_int64, [C.c_void_p, C.c_size_t]),
    "rtjx_region_get": (C.c_int64, [C.c_void_p, C.c_size_t, C.POINTER(Junction), C.c_size_t]),
    "rtjx_unique_junctions": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.c_size_t]),
    "rtjx_unique_count": (C.c_int64, [C.c_void_p]),
    "rtjx_unique_get": (C.c_int64, [C.c_void_p, C.POINTER(Junction), C.POINTER(C.c_uint32), C.c_size_t]),
    "rtjx_unique_regions": (C.c_int64, [C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32), C.c_size_t]),
    "rtjx_scan_batch": (C.c_int, [C.c_void_p, C.POINTER(Batch), C.c_int, C.c_void_p]),
    "rtjx_add": (C.c_int, [C.c_void_p, C.POINTER(Candidate), C.c_size_t]),
    "rtjx_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rtjx_count": (C.c_int64, [C.c_void_p]),
    "rtjx_get": (C.c_int64, [C.c_void_p, C.POINTER(Junction), C.c_size_t]),
    "rtjx_write_bed12": (C.c_int, [C.c_void_p, C.c_int]),
    "rtjx_intern_barcode": (C.c_int64, [C.c_void_p, C.c_char_p]),
    "rtjx_write_barcodes": (C.c_int, [C.c_void_p, C.c_int]),
    "rtjx_barcode_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "rtjx_barcode_name": (C.c_char_p, [C.c_void_p, C.c_uint32]),
    "rtjx_load_barcodes": (C.c_int64, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "rtjx_import": (C.c_int, [C.c_void_p, C.POINTER(Junction), C.c_size_t]),
    "rtjx_contig": (C.c_char_p, [C.c_void_p, C.c_int32]),
    "rtjx_n_contigs": (C.c_int32, [C.c_void_p]),
    "rtjx_intern_contig": (C.c_int32, [C.c_void_p, C.c_char_p]),
    "rtjx_plan_shards": (C.c_int32, [C.c_char_p, C.c_int32, C.POINTER(C.c_int32), C.c_size_t]),
    "rtjx_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "rtjx_reset_stats": (None, [C.c_void_p]),
    "rtjx_clear": (C.c_int, [C.c_void_p]),
    "rtjx_load_batch": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rtjx_inflate_file": (C.c_int, [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "rtjx_stage_bam": (C.c_int, [C.c_void_p]),
    "rtjx_comm_unique_id": (C.c_int, [C.c_void_p]),
    "rtjx_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    "rtjx_comm_destroy": (None, []),
    "rtjx_gather": (C.c_int, [C.c_void_p, C.c_int]),
    "rtjx_annotate_params_default": (None, [C.POINTER(AnnotateParams)]),
    "rtjx_annotate": (C.c_int, [C.POINTER(AnnotateParams), C.c_int, C.POINTER(C.c_uint64), C.c_char_p, C.c_size_t]),
    "rtjx_last_error": (C.c_char_p, [C.c_void_p]),
    "rtjx_strerror": (C.c_char_p, [C.c_int]),
    "rtjx_version": (C.c_char_p, []),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)      # AttributeError here = the library does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args
