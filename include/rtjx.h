/* include/rtjx.h — C ABI of libregtools_jx.so, the B200-native drop-in for the
 * `regtools junctions extract` hot path.
 *
 * The reference has no FFI; its boundary for this path is the C++ class JunctionsExtractor
 * (/root/reference/src/junctions/junctions_extractor.h:149-248) driven by
 * src/junctions/junctions_main.cc:45-59 and by
 * src/cis-splice-effects/cis_splice_effects_identifier.cc:288-290.  Every entry point below
 * names the reference member it replaces.  The C++ shim with the reference's own class and
 * method names lives in regtools_b200/csrc/junctions_extractor.h; INTEGRATION.md shows the
 * binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 (RTJX_OK) or a negative
 * rtjx_status and never throws; the message for the last failure of a handle is available
 * from rtjx_last_error().  A handle is single-owner (one caller thread at a time); several
 * handles may coexist.  There is NO CPU fallback: without a usable CUDA device every compute
 * entry point fails with RTJX_E_CUDA.
 */
#ifndef RTJX_H
#define RTJX_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    RTJX_OK            = 0,
    RTJX_E_ARG         = -1,  /* bad argument / parameter combination                              */
    RTJX_E_OPEN_BAM    = -2,  /* "Unable to open BAM/SAM file."          junctions_extractor.cc:505 */
    RTJX_E_OPEN_INDEX  = -3,  /* "Unable to open BAM/SAM index. ..."     junctions_extractor.cc:510 */
    RTJX_E_REGION      = -4,  /* "Unable to iterate to region within BAM." junctions_extractor.cc:521 */
    RTJX_E_CUDA        = -5,  /* no device / CUDA runtime failure                                   */
    RTJX_E_UNSUPPORTED = -6,  /* CRAM/SAM input, compressed FASTA, -b combined with shards/regions  */
    RTJX_E_NOMEM       = -7,
    RTJX_E_STATE       = -8,  /* call order violated                                                */
    RTJX_E_IO          = -9   /* also "Unable to extract FASTA sequence ..."  junctions_extractor.cc:553 */
} rtjx_status;

typedef struct rtjx_handle rtjx_t;

/* Parameters = the private members the reference sets in parse_options() / the 8-arg ctor
 * (junctions_extractor.h:151-181,199-205; CLI flags junctions_extractor.cc:46-110). */
typedef struct {
    uint32_t    struct_size;      /* = sizeof(rtjx_params), set by rtjx_params_default         */
    const char* bam;              /* bam_      positional 1; may be NULL for add-only handles  */
    const char* region;           /* region_   -r, default "."                                 */
    const char* strand_tag;       /* strand_tag_ -t, default "XS" (first two chars used)       */
    const char* fasta;            /* ref_      positional 2: uncompressed FASTA; strand from the
                                     intron motif first, -s only for '?' (junctions_extractor.cc:325-359) */
    const char* barcode_out;      /* output_barcodes_file_ -b: non-NULL switches the handle to the single-cell mode
                                     (junctions_extractor.cc:203-215,362-374,393-395); the engine does not open the
                                     path itself, the caller writes it with rtjx_write_barcodes       */
    int32_t     strandness;       /* -s: 0 XS, 1 RF, 2 FR, 3 intron-motif(no FASTA => as FR)   */
    uint32_t    min_anchor;       /* -a, default 8                                             */
    uint32_t    min_intron;       /* -m, default 70                                            */
    uint32_t    max_intron;       /* -M, default 500000                                        */
    int32_t     device;           /* CUDA ordinal; -1 = host-only table handle (import/print)  */
    int32_t     n_threads;        /* host BGZF inflate workers; 0 = all cores                  */
    uint32_t    batch_reads;      /* reads per device batch; 0 = default                       */
    uint32_t    table_log2;       /* initial junction hash capacity = 2^table_log2; 0 = default*/
    int32_t     shard_rank;       /* contig shard of this handle (multi-GPU), default 0        */
    int32_t     shard_world;      /* number of shards, default 1                               */
    int32_t     inflate_mode;     /* 0 auto, 1 host zlib workers, 2 device inflate kernel      */
    int32_t     profile;          /* 1: bracket every kernel with CUDA events (rtjx_get_stats) */
    int32_t     scan_variant;     /* cigar_scan kernel: 0 / 5 = block per tile (default), 8 = warp-pipelined persistent   */
    int32_t     scan_cfg;         /* configuration of the warp-pipelined kernel (ring depth / warps per SM), 0 = default */
    const char* barcode_tag;      /* barcode_tag_ (junctions_extractor.h:181), NULL = "CB"; the reference has no flag for it */
} rtjx_params;

/* One merged junction; same fields as the reference's struct Junction
 * (junctions_extractor.h:39-112) minus the heap strings. */
typedef struct {
    int32_t  tid;                 /* index into the BAM header's targets (or interned name)   */
    uint32_t start, end;          /* BED::start/end of the intron                             */
    uint32_t thick_start, thick_end;
    uint32_t read_count;
    uint32_t name_index;          /* N of "JUNC%08d"                                          */
    uint8_t  strand;              /* printed strand char                                      */
    uint8_t  left_ok, right_ok;   /* has_left_min_anchor / has_right_min_anchor               */
    uint8_t  pad;
    uint64_t first_ord;           /* ordinal of the first supporting N op (naming order)      */
} rtjx_junction;

/* One candidate as the reference passes to add_junction(Junction) (before junction_qc). */
typedef struct {
    int32_t  tid;
    uint32_t start, end, thick_start, thick_end;
    uint8_t  strand;              /* strand char: '+', '-', anything else                     */
    uint8_t  pad[3];
} rtjx_candidate;

/* A SoA batch of alignments, the unit handed to the cigar_scan kernel.
 * Algorithmic bytes: 16*n_reads + 4*n_ops read.  meta = flag<<16 | mapq<<8 | strand_byte, where
 * strand_byte is the value of the strand tag if its type is 'A', else 0.  cig_off has
 * n_reads+1 entries indexing `cigar` (uint32 len<<4|op, htslib/sam.h:75-83). */
typedef struct {
    uint32_t        n_reads;
    uint32_t        n_ops;
    uint64_t        first_ordinal; /* ordinal of read 0 in iteration order                    */
    uint32_t        n_junction_ops;/* number of N ops in `cigar` if known, else 0: the engine
                                      then synchronises once after cigar_scan to size the merge */
    uint32_t        reserved;
    const int32_t*  tid;
    const int32_t*  pos;
    const uint32_t* meta;
    const uint32_t* cig_off;
    const uint32_t* cigar;
    const uint32_t* bc;            /* `-b` handles only (else ignored, may be NULL): per alignment the id rtjx_intern_barcode
                                      returned for its barcode (set_junction_barcode, junctions_extractor.cc:362-374: the CB:Z
                                      value, "?" when absent); read for n_cigar > 1 alignments                          */
} rtjx_batch;

typedef struct {
    uint64_t reads;               /* alignments iterated (incl. n_cigar<=1)                   */
    uint64_t cigar_ops;
    uint64_t candidates;          /* N ops emitted by cigar_scan                              */
    uint64_t batches;
    uint64_t kernel_launches;     /* launches of OUR kernels                                  */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t bgzf_blocks, compressed_bytes, inflated_bytes;
    double   scan_ms, merge_ms, finalize_ms, inflate_kernel_ms;  /* device time, profile=1    */
    double   host_inflate_s, host_parse_s, host_wait_s, total_s; /* wall, summed over threads */
    uint32_t table_slots, table_grows;
} rtjx_stats;

#define RTJX_LOC_HOST   0
#define RTJX_LOC_DEVICE 1

void        rtjx_params_default(rtjx_params* p);

/* JunctionsExtractor::JunctionsExtractor (junctions_extractor.h:184-205) */
int         rtjx_create(const rtjx_params* p, rtjx_t** out);
void        rtjx_destroy(rtjx_t* h);

/* JunctionsExtractor::identify_junctions_from_BAM (junctions_extractor.cc:500-535):
 * open + index + header + iterate region + per-read CIGAR walk + merge. */
int         rtjx_run(rtjx_t* h);

/* The second caller of the path, batched (src/cis-splice-effects/cis_splice_effects_identifier.cc:267-311 builds one
 * JunctionsExtractor per variant region on the same BAM: `JunctionsExtractor je1(bam_, variant_region, ...);
 * je1.identify_junctions_from_BAM(); je1.get_all_junctions()`).  rtjx_run_regions computes the table of EVERY region
 * in one pass over the file; region i then answers exactly what a fresh handle created with region = regions[i]
 * would: own JUNC numbering, sorted by compare_junctions, no anchor filter.  The handle must have been created with
 * region "." (parameters, strand mode and FASTA apply to all regions); its own table is left empty. */
int         rtjx_run_regions(rtjx_t* h, const char* const* regions, size_t n_regions);
int64_t     rtjx_region_count(rtjx_t* h, size_t region);
int64_t     rtjx_region_get(rtjx_t* h, size_t region, rtjx_junction* out, size_t cap);

/* The rest of that caller's loop (cis_splice_effects_identifier.cc:292-299), on the tables of the last rtjx_run_regions:
 * region i stands for variant i with the window [win_start[i], win_end[i]] (v1.cis_effect_start / cis_effect_end, both
 * inclusive).  Every junction of region i whose start or end lies inside the window (:294-295) is inserted, regions in order
 * and junctions in table order, into the reference's `set<Junction> unique_junctions_` and `map<Junction, set<variant>>
 * junction_to_variant_`.  Both compare through the implicit Junction -> AnnotatedJunction conversion
 * (junctions_annotator.h:155-177): contig name (std::string <), start, end — strand-blind — and the first insert wins.
 * rtjx_unique_get returns the set in its iteration order (the record of the winning insert, `first_region` = its variant);
 * rtjx_unique_regions(i) the variants of unique junction i in ascending order.  Counts are returned, negative = error. */
int         rtjx_unique_junctions(rtjx_t* h, const uint32_t* win_start, const uint32_t* win_end, size_t n_regions);
int64_t     rtjx_unique_count(rtjx_t* h);
int64_t     rtjx_unique_get(rtjx_t* h, rtjx_junction* out, uint32_t* first_region, size_t cap);
int64_t     rtjx_unique_regions(rtjx_t* h, size_t i, uint32_t* out, size_t cap);

/* parse_alignment_into_junctions over a prepared batch (junctions_extractor.cc:377-497):
 * launches cigar_scan + junction_merge on `stream` (a cudaStream_t, NULL = default stream).
 * location = RTJX_LOC_HOST: arrays are host memory and are copied in; RTJX_LOC_DEVICE: the
 * arrays already live in HBM (16-byte aligned) and are used in place. */
int         rtjx_scan_batch(rtjx_t* h, const rtjx_batch* b, int location, void* stream);

/* JunctionsExtractor::add_junction (junctions_extractor.cc:174-235) for n candidates, in order. */
int         rtjx_add(rtjx_t* h, const rtjx_candidate* c, size_t n);

/* Makes the merged table current (compaction, first-seen ranking, sort); implied by the getters. */
int         rtjx_finalize(rtjx_t* h, void* stream);

/* junctions_.size() after QC (get_new_junction_name, junctions_extractor.cc:152-157). */
int64_t     rtjx_count(rtjx_t* h);
/* get_all_junctions (junctions_extractor.cc:238-246): ALL junctions, sorted by
 * compare_junctions (junctions_extractor.h:117-140), no anchor filter.  Returns the total
 * number; fills at most cap. */
int64_t     rtjx_get(rtjx_t* h, rtjx_junction* out, size_t cap);
/* print_all_junctions (junctions_extractor.cc:249-280): BED12 of the anchor-filtered, sorted
 * junctions to a file descriptor. */
int         rtjx_write_bed12(rtjx_t* h, int fd);
/* `-b` single-cell mode (handle created with rtjx_params.barcode_out != NULL; fed by rtjx_run, or by rtjx_scan_batch
 * with rtjx_batch.bc; rtjx_add carries no barcode and refuses).
 * rtjx_write_barcodes = Junction::print_barcodes (junctions_extractor.h:99-111) for every junction that
 * print_all_junctions prints (junctions_extractor.cc:267-273), in the same order: "<n>\t<bc>:<count>,...\n" with the
 * barcodes in the iteration order of the reference's std::unordered_map.
 * rtjx_barcode_stats: distinct values of the barcode tag seen (incl. "?") and the number of n_cigar > 1 alignments
 * without the tag (= the reference's "WARNING: No CB tag found ..." lines, junctions_extractor.cc:371).
 * rtjx_barcode_name: dictionary entry `id` (NULL past the end).
 * rtjx_load_barcodes: host feeder only, like rtjx_load_batch: the per-alignment dictionary ids (0 for n_cigar <= 1) of
 * the handle's region in iteration order; returns the number of alignments, fills at most cap. */
/* Batch-level `-b` (rtjx_scan_batch on a -b handle): registers a barcode string and returns its dictionary id (>= 0), the
 * value to put into rtjx_batch.bc; a repeated string returns the same id. */
int64_t     rtjx_intern_barcode(rtjx_t* h, const char* barcode);
int         rtjx_write_barcodes(rtjx_t* h, int fd);
int         rtjx_barcode_stats(rtjx_t* h, uint64_t* n_barcodes, uint64_t* n_missing);
const char* rtjx_barcode_name(rtjx_t* h, uint32_t id);
int64_t     rtjx_load_barcodes(rtjx_t* h, uint32_t* ids, size_t cap);

/* Entries of other shards (disjoint contigs) are appended; names are then ranked by
 * (tid, first_ord), which equals BAM order for a coordinate-sorted file. */
int         rtjx_import(rtjx_t* h, const rtjx_junction* j, size_t n);

/* header->target_name[tid] (junctions_extractor.cc:384) / interning for add-only handles. */
const char* rtjx_contig(rtjx_t* h, int32_t tid);
int32_t     rtjx_n_contigs(rtjx_t* h);
int32_t     rtjx_intern_contig(rtjx_t* h, const char* name);

/* Contig -> shard assignment (LPT over compressed bytes per contig from the BAI pseudo-bin,
 * hts.c:1092); assign[tid] in [0,world).  Returns n_contigs or a negative status. */
int32_t     rtjx_plan_shards(const char* bam, int32_t world, int32_t* assign, size_t cap);

int         rtjx_get_stats(rtjx_t* h, rtjx_stats* out);
void        rtjx_reset_stats(rtjx_t* h);
/* Drops all junctions but keeps device buffers (re-run on the same handle). */
int         rtjx_clear(rtjx_t* h);

/* Host feeder only (BGZF inflate + BAM record split into SoA): fills caller-allocated arrays
 * for kernel-level tests and benches.  Pass NULL arrays to size: *n_reads / *n_ops are set. */
int         rtjx_load_batch(rtjx_t* h, uint64_t* n_reads, uint64_t* n_ops, int32_t* tid, int32_t* pos,
                            uint32_t* meta, uint32_t* cig_off, uint32_t* cigar);

/* Device BGZF inflate of the handle's BAM, from the first block, at most max_blocks blocks (0 = all):
 * the inflated byte stream is copied to `out` (host, cap bytes); *out_len receives its length.
 * Test hook for the inflate kernel (replaces bgzf.c:292-316 for whole-file runs). */
int         rtjx_inflate_file(rtjx_t* h, uint64_t max_blocks, void* out, uint64_t cap, uint64_t* out_len);
/* Copies the handle's BAM — the compressed file, byte for byte — into device memory.  Later rtjx_run / rtjx_run_regions calls on
 * this handle read it from there (no host staging, no H2D): the "input already resident in HBM" configuration of bench.py.
 * The reference has no counterpart (it reads through bgzf_read, bgzf.c:548-577, on every run); results are unchanged. */
int         rtjx_stage_bam(rtjx_t* h);

/* ---- multi-GPU: the path's one exchange -------------------------------------------------------------
 * One process per GPU; every rank runs rtjx_run on its contig shard (rtjx_params.shard_rank / shard_world).  The junction
 * keys of different contigs are disjoint, so the merge is a concatenation + the global first-seen ranking + the sort.
 * rtjx_gather does it with NCCL from inside the library: counts all-gathered, every rank's compacted table sent from HBM to
 * the root's HBM (ncclSend / ncclRecv, exact sizes), names ranked and the table sorted on the root's GPU, one D2H.
 * Afterwards the ROOT's handle serves the merged table (rtjx_count / rtjx_get / rtjx_write_bed12); the other ranks keep their
 * own shard.  The communicator is process-wide: rank 0 calls rtjx_comm_unique_id and hands the bytes to the other ranks by
 * whatever launcher it has (MPI, torch.distributed, a file); every rank then calls rtjx_comm_init once.  NCCL is loaded with
 * dlopen at that moment, not before.  The reference has no counterpart (single process, single thread). */
#define RTJX_COMM_ID_BYTES 128
int         rtjx_comm_unique_id(void* id /* RTJX_COMM_ID_BYTES */);
int         rtjx_comm_init(rtjx_t* h, const void* id, int rank, int world);
void        rtjx_comm_destroy(void);
int         rtjx_gather(rtjx_t* h, int root);

/* ---- `regtools junctions annotate` (the downstream consumer of the BED12) -------------------------
 * One call = junctions_annotate (src/junctions/junctions_main.cc:61-92): load the GTF (gtf_parser.cc), read the BED12
 * junctions (BedFile), and for every line adjust_junction_ends + get_splice_site + annotate_junction_with_gtf
 * (junctions_annotator.cc:66-81,94-114,367-388 -> overlap_ps/overlap_ns :128-311) + AnnotatedJunction::print
 * (junctions_annotator.h:86-121).  The per-junction work runs on the device (one thread per junction over flat GTF
 * arrays and the genome in HBM); parsing and the TSV text are host work.  Header + one line per junction go to out_fd.
 * Errors keep the reference's order: a GTF problem fails before anything is written; a junction on a contig the FASTA
 * lacks ("Unable to extract FASTA sequence for position ...", RTJX_E_IO) or a line that is not BED12 ends the run after
 * the lines before it were written.  err (optional) receives the message. */
typedef struct {
    uint32_t    struct_size;          /* = sizeof(rtjx_annotate_params), set by rtjx_annotate_params_default        */
    const char* junctions_bed;        /* junctions_.bedFile   positional 1                                          */
    const char* fasta;                /* ref_                 positional 2 (uncompressed FASTA)                     */
    const char* gtf;                  /* gtf_                 positional 3                                          */
    int32_t     include_single_exon;  /* -S: skip_single_exon_genes_ = false (junctions_annotator.cc:411-413)       */
    int32_t     device;               /* CUDA ordinal                                                               */
    int32_t     chatter_fd;           /* >= 0: the reference's stderr lines ("position = ...", "Annotated N lines.") */
    const char* out_path;             /* output_file_ -o: used when out_fd < 0; opened only after the GTF and the junctions
                                         file were read, as set_ofstream_object is (junctions_main.cc:68-71)          */
} rtjx_annotate_params;
void        rtjx_annotate_params_default(rtjx_annotate_params* p);
int         rtjx_annotate(const rtjx_annotate_params* p, int out_fd, uint64_t* n_lines, char* err, size_t err_cap);

const char* rtjx_last_error(const rtjx_t* h);
const char* rtjx_strerror(int status);
const char* rtjx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RTJX_H */
