// regtools_b200/csrc/jx_device.cuh — device-side data layout shared by the kernels and the engine.
//
// Everything on the hot path is 32/64-bit integer work bounded by HBM bandwidth; there is no
// GEMM-shaped step and therefore no tensor-core use.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rtjx {

// One raw candidate = one N op of one alignment, written by cigar_scan, read by junction_merge.
// 32 bytes, 16-byte aligned so a thread moves it as two 128-bit accesses.
struct alignas(16) Cand {
    uint32_t start, end;          // intron [start,end)                      (A.2: start_k, end_k)
    uint32_t ts, te;              // thick_start / thick_end of THIS read
    uint64_t ord;                 // read_ordinal << 16 | k   (k = index of the N op in the CIGAR)
    int32_t  tid;
    uint32_t strand;              // strand char of the read in the low byte
};
static_assert(sizeof(Cand) == 32, "Cand must be 32 bytes");

// One slot of the device-wide open-addressed junction table (linear probing).
// The 128-bit key is claimed with a single ATOMG.CAS.128; an all-zero slot is empty, so the
// table is (re)initialised with cudaMemsetAsync.  min-reductions are stored complemented so that
// zero is the identity of every field:
//   count : atomicAdd        nts   : atomicMax(~thick_start)  te : atomicMax(thick_end)
//   lr    : atomicOr (bit0 = left anchor ok, bit1 = right anchor ok)
//   nfirst: atomicMax(~first_ord)   -> name rank  (first-seen,  junctions_extractor.cc:152-157,198)
//   last  : atomicMax(read_ord<<8|strand_char) -> printed strand (last writer, :229), proxy 2 only
struct alignas(16) Slot {
    unsigned long long klo;       // start << 32 | end
    unsigned long long khi;       // (tid + 1) << 2 | proxy      (non-zero for every real key)
    uint32_t count, nts, te, lr;
    unsigned long long nfirst, last;
};
static_assert(sizeof(Slot) == 48, "Slot must be 48 bytes");

// Compacted table entry produced by finalize (same layout as rtjx_junction in include/rtjx.h).
// In the batched variant-region mode the table key carries the region in khi bits 34.. (khi = (region + 1) << 34 |
// (tid + 1) << 2 | proxy) and finalize works on this wider record.
struct OutJunctionR;
struct alignas(8) OutJunction {
    int32_t  tid;
    uint32_t start, end, ts, te, count, name_index;
    uint8_t  strand, left_ok, right_ok, pad;
    uint64_t first_ord;
};
static_assert(sizeof(OutJunction) == 40, "OutJunction must be 40 bytes");
struct alignas(8) OutJunctionR { OutJunction j; uint32_t region; uint32_t pad; };
static_assert(sizeof(OutJunctionR) == 48, "OutJunctionR must be 48 bytes");

// Batched variant regions (second caller of the path, cis_splice_effects_identifier.cc:267-311: one extractor per
// variant region on the same BAM).  Sorted by (tid, beg); [beg, end) is the 0-based half-open interval hts_itr_query
// iterates (hts.c:1897-1922): an alignment belongs to a region iff tid matches, pos < end and endpos > beg
// (hts.c:1941-1963).  n == 0: mode off.
struct VariantRegions {
    const int32_t* tid; const int32_t* beg; const int32_t* end;
    uint32_t n;
    uint32_t max_len;             // max(end - beg), bounds the backward search
    uint32_t tag;                 // 1: candidates carry their region's index (rtjx_run_regions); 0: the regions only FILTER (a `-r` run on the device feeder)
};

struct ScanParams {
    int32_t  strandness;          // 0 XS, 1 RF, 2/3 FR
    uint32_t min_anchor, min_intron, max_intron;
    int32_t  variant;             // cigar_scan kernel: 5 = block per tile (default), 8 = warp-pipelined persistent
    int32_t  cfg;                 // tile configuration of the variant (0 = production)
    // intron-motif strand mode (a FASTA was given): the genome as one byte per base in HBM.  NULL = off.
    const uint8_t*            genome;
    const unsigned long long* g_off;      // per BAM tid: start of the contig in `genome`
    const unsigned long long* g_len;      // per BAM tid: length, ~0ull = the FASTA has no such sequence
    uint32_t                  g_n;        // entries of g_off / g_len
    VariantRegions            vr;         // batched variant regions; candidates then carry (region index + 1) << 8 in `strand`
};

struct BatchView {
    uint32_t n_reads, n_ops;
    uint64_t first_ordinal;
    const int32_t*  tid;
    const int32_t*  pos;
    const uint32_t* meta;
    const uint32_t* cig_off;
    const uint32_t* cigar;
    // `-b` single-cell mode (set_junction_barcode, junctions_extractor.cc:362-374): dictionary id of the alignment's barcode
    // (meaningful for n_cigar > 1 only).  NULL = mode off.  Candidates then carry (id + 1) << 8 in `strand`, exactly where
    // the variant-region mode puts its region, so the (junction, barcode) pair is the table key.
    const uint32_t* bc = nullptr;
};

// The device-wide junction table as the kernels see it.
struct TableRef {
    Slot*     slots;
    uint32_t  mask;               // slots - 1 (power of two)
    uint32_t* slot_list;          // slot index of the i-th distinct junction (i < CTR_NUNIQUE)
    uint32_t  list_cap;
};

// Launchers (kernels.cu).  All are asynchronous on `stream`.
// Candidate slots cigar_scan may reserve beyond the N ops of a batch (the warp-pipelined kernel reserves in chunks and pads
// the tail of every warp's last chunk with tid = -1 entries): size the candidate buffer, and junction_merge's bound, as
// N ops + this.
uint32_t cigar_scan_cand_slack();
void launch_cigar_scan(const BatchView& b, const ScanParams& p, Cand* cands, uint32_t cand_cap,
                       uint32_t* d_counters /* [0]=n_cand, [1]=overflowed */, cudaStream_t stream);
void launch_junction_merge(const Cand* cands, const uint32_t* d_n_cand, uint32_t n_cand_bound, const ScanParams& p, const TableRef& tb,
                           Slot* spill, uint32_t spill_cap, uint32_t* d_counters /* [2]=n_unique,[3]=n_spill */, cudaStream_t stream);
void launch_table_rehash(const Slot* old_table, uint32_t old_slots, const TableRef& tb, uint32_t* d_counters,
                         cudaStream_t stream);
void launch_table_clear(const TableRef& tb, const uint32_t* d_n_unique, uint32_t n_bound, cudaStream_t stream);
// `-b` mode: the handle's table is keyed (junction, barcode).  Every field of a slot is an associative reduction (add, max,
// max, or, max, max), so the junction-level table of add_junction is the fold of the pair table over the barcode bits of the
// key: the n occupied slots of `src` are upserted into `dst` with khi bits 34.. cleared.  dst_counters: CTR_* of `dst`.
void launch_table_fold(const TableRef& src, uint32_t n, const TableRef& dst, uint32_t* dst_counters, cudaStream_t stream);
// (junction, barcode) pairs of a `-b` table, sorted by (tid, start, end, proxy, first_ord): one run per junction, barcodes in
// first-seen order.  Workspace: finalize_sort_regions_workspace_bytes(n).
void launch_sort_barcode_pairs(OutJunctionR* entries, uint32_t n, void* workspace, size_t workspace_bytes, cudaStream_t stream);
void launch_table_compact(const TableRef& tb, uint32_t n, OutJunction* out, cudaStream_t stream);
// variant-region mode: compaction into OutJunctionR, names ranked per region, sorted by (region, contig, ts, te, name)
void launch_table_compact_regions(const TableRef& tb, uint32_t n, OutJunctionR* out, cudaStream_t stream);
size_t finalize_sort_regions_workspace_bytes(uint32_t n);
void launch_finalize_sort_regions(OutJunctionR* entries, uint32_t n, const uint32_t* contig_rank, uint32_t n_contigs,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream);
// ranks entries by first_ord (name_index) and sorts them in place by (contig_rank[tid], ts, te, name_index)
size_t finalize_sort_workspace_bytes(uint32_t n);
// rank_by_contig: the entries come from several contig shards (rtjx_gather): first appearance = (contig, first_ord) order
void launch_finalize_sort(OutJunction* entries, uint32_t n, const uint32_t* contig_rank, uint32_t n_contigs,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream, bool rank_by_contig = false);

// BGZF inflate on the device (inflate.cu).  `blocks` is an array of BgzfBlockDesc in device memory.
struct BgzfBlockDesc { uint32_t in_off, in_len, out_off, out_len; };
// scratch: bgzf_inflate_scratch_bytes(n_blocks) bytes of device memory for the lane-per-stream decoder's match lists
// (NULL: only the warp-per-block kernel is used).
size_t bgzf_inflate_scratch_bytes(uint32_t n_blocks);
void launch_bgzf_inflate(const uint8_t* comp, const void* blocks, uint32_t n_blocks, uint8_t* out, uint32_t* status,
                         void* scratch, cudaStream_t stream);

void launch_inflate_status_reduce(const uint32_t* status, uint32_t n_blocks, uint32_t* flags, cudaStream_t stream);

// Device-side BAM record split (device_feed.cu).
struct FeedState {
    long long bad_offset;         // smallest offset of a malformed record (LLONG_MAX: none)
    long long next_carry_from;    // offset of the first byte that belongs to the next chunk
    uint32_t carry_len;           // bytes carried in front of the next chunk's data
    uint32_t flags;
    uint32_t n_rec, n_ops, n_junction_ops;
    uint32_t reached_limit;       // the range end (contig shard) lies in this chunk
    // alignments appended to the SoA accumulator since the last cigar_scan (survive feed_reset; cleared by feed_acc_reset)
    uint32_t acc_rec, acc_ops, acc_jops, pad;
};
enum { FEED_FLAG_SEED_MISS = 1, FEED_FLAG_CAPACITY = 2, FEED_FLAG_CARRY_TOO_BIG = 4, FEED_FLAG_INFLATE = 8 };
size_t feed_scan_workspace_bytes(uint32_t n);
void launch_feed_reset(FeedState* state, int keep_carry, cudaStream_t stream);
// seeds[1 .. n_blocks) = first record start inside BGZF block i found by plausibility checks (verified by record_walk's chain);
// seeds[0] is left to the caller.  blocks: BgzfBlockDesc[n_blocks] in device memory.
void launch_block_seeds(const uint8_t* data, int64_t data_len, const void* blocks, uint32_t n_blocks, int32_t n_ref, int64_t* seeds,
                        cudaStream_t stream);
void launch_record_walk(const uint8_t* data, int64_t data_len, int64_t limit, const int64_t* seeds, const uint32_t* seg_base,
                        uint32_t n_seg, int use_carry, FeedState* state, int32_t* rec_off, uint32_t* seg_cnt,
                        cudaStream_t stream);
void launch_record_gather(const uint8_t* data, const int32_t* rec_off, const uint32_t* seg_base, uint32_t* seg_cnt,
                          uint32_t* seg_scan, uint32_t n_seg, uint32_t cap_total, FeedState* state, int32_t* dense,
                          uint32_t* ncig, uint32_t* ncig_scan, void* ws, size_t ws_bytes, cudaStream_t stream);
// The group's alignments are appended to the accumulator arrays at state->acc_rec / state->acc_ops (rec_cap / cigar_cap: their sizes).
void launch_record_extract(const uint8_t* data, const int32_t* dense, const uint32_t* ncig_scan, uint32_t cap_total,
                           FeedState* state, int32_t n_ref, int xs_mode, uint32_t tag0, uint32_t tag1, int32_t* o_tid,
                           int32_t* o_pos, uint32_t* o_meta, uint32_t* o_off, uint32_t* o_cigar, uint32_t rec_cap, uint32_t cigar_cap,
                           cudaStream_t stream);
// carry: the unfinished record feed_finish left right-aligned in front of `carry_end` is copied in front of `data`
void launch_feed_carry_in(const uint8_t* carry_end, uint8_t* data, const FeedState* state, cudaStream_t stream);
void launch_feed_acc_reset(FeedState* state, cudaStream_t stream);
void launch_feed_finish(const uint8_t* data, int64_t data_len, uint8_t* next_data, uint32_t headroom, const uint32_t* seg_scan,
                        uint32_t n_seg, const uint32_t* ncig_scan, FeedState* state, cudaStream_t stream);

// counters layout in d_counters (uint32 each)
enum { CTR_NCAND = 0, CTR_CAND_OVERFLOW = 1, CTR_NUNIQUE = 2, CTR_NSPILL = 3, CTR_NOUT = 4, CTR_UNUSED5 = 5,
       CTR_TOTAL_CAND64 = 6 /* 64-bit, two words */, CTR_GENOME_MISS = 8 /* 1 + tid of a contig the FASTA lacks */,
       CTR_GENOME_MISS_POS = 9 /* start of one such junction */, CTR_COUNT = 10 };

}  // namespace rtjx
