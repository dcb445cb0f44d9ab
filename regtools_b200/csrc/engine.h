// regtools_b200/csrc/engine.h — per-handle state behind the C ABI (include/rtjx.h).
//
// Owns the device-side junction table, the candidate buffer, the pinned/HBM batch rings and the
// host copy of the finalized table.  Mirrors the life cycle of the reference's JunctionsExtractor
// object (/root/reference/src/junctions/junctions_extractor.h:149-248).
#pragma once
#include "../../include/rtjx.h"
#include "bam_feeder.h"
#include "fasta.h"
#include "jx_device.cuh"

#include <nvtx3/nvToolsExt.h>

#include <memory>
#include <string>
#include <vector>

namespace rtjx {

// NVTX range over a scope: the stages of a run show up by name in Nsight Systems / ncu --nvtx (SURVEY §5 tracing row)
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// process-wide NCCL communicator (exchange.cc)
int comm_unique_id(void* id, std::string* err);
int comm_init(const void* id, int rank, int world, int device, std::string* err);
void comm_destroy();

class Engine {
public:
    explicit Engine(const rtjx_params& p);
    ~Engine();

    int run();                                   // identify_junctions_from_BAM
    // cis_splice_effects_identifier.cc:267-311 batched: every region's table from ONE pass over the BAM
    int run_regions(const char* const* regions, size_t n);
    int64_t region_count(size_t i);
    int64_t region_get(size_t i, rtjx_junction* out, size_t cap);
    // cis_splice_effects_identifier.cc:292-299 on the tables of the last run_regions: the unique-junction set + junction -> variants
    int unique_build(const uint32_t* win_start, const uint32_t* win_end, size_t n);
    int64_t unique_count();
    int64_t unique_get(rtjx_junction* out, uint32_t* first_region, size_t cap);
    int64_t unique_regions(size_t i, uint32_t* out, size_t cap);
    int scan_batch(const rtjx_batch& b, int location, cudaStream_t stream);
    int add(const rtjx_candidate* c, size_t n);
    int finalize(cudaStream_t stream);
    int64_t count();
    int64_t get(rtjx_junction* out, size_t cap);
    int write_bed12(int fd);
    // `-b` single-cell mode (junctions_extractor.cc:203-215,362-374; print_barcodes junctions_extractor.h:99-111)
    bool barcode_mode() const { return bc_mode_; }
    int write_barcodes(int fd);
    int64_t intern_barcode(const char* s);
    int barcode_stats(uint64_t* n_barcodes, uint64_t* n_missing);
    const char* barcode_name(uint32_t id);
    int64_t load_barcodes(uint32_t* ids, size_t cap);          // host feeder only: per-alignment dictionary ids
    int import(const rtjx_junction* j, size_t n);
    int clear();
    int inflate_file(uint64_t max_blocks, void* out, uint64_t cap, uint64_t* out_len);
    int gather(int root);                        // rtjx_gather: NCCL exchange of the shard tables, merged table on the root (exchange.cc)
    std::vector<uint32_t> contig_rank_table() const;
    int stage_file();                            // rtjx_stage_bam: compressed BAM -> HBM; later runs read it from there (device_run.cc)
    int load_batch(uint64_t* n_reads, uint64_t* n_ops, int32_t* tid, int32_t* pos, uint32_t* meta,
                   uint32_t* cig_off, uint32_t* cigar);

    const char* contig(int32_t tid);
    int32_t n_contigs() const { return (int32_t)contigs_.size(); }
    int32_t intern_contig(const char* name);

    void get_stats(rtjx_stats* out);
    void reset_stats();
    const char* last_error() const { return err_.c_str(); }
    int fail(int status, const std::string& msg) { err_ = msg; return status; }
    bool host_only() const { return prm_.device < 0; }
    int device() const { return prm_.device; }

private:
    friend struct EngineSink;
    int ensure_device();
    int run_impl();
    int ensure_genome();                         // FASTA -> HBM (once) + per-contig map (whenever the contig list changed)
    int ensure_cands(uint32_t n);
    int ensure_table(uint32_t incoming_bound, cudaStream_t stream);
    int process_device_batch(const BatchView& v, uint32_t cand_bound, cudaStream_t stream);
    int sync_counters(cudaStream_t stream);
    int open_bam(std::unique_ptr<BamFile>* bam, BaiIndex* idx, IterSpec* spec);
    int run_device(const BamFile& bam, const BaiIndex& idx, const IterSpec& spec);   // device_run.cc; +1 = declined
    void host_rank_and_sort();
    void resolve_profile_events();
    ScanParams scan_params() const;

    rtjx_params prm_;
    std::string bam_path_, region_, tag_, fasta_path_;
    std::string err_;
    std::vector<std::string> contigs_;
    std::string unknown_contig_;

    // device state
    bool dev_ready_ = false;
    cudaStream_t stream_ = nullptr, copy_stream_ = nullptr;
    uint32_t* d_counters_ = nullptr;
    uint32_t* h_counters_ = nullptr;            // pinned mirror
    Slot* d_table_ = nullptr; uint32_t table_slots_ = 0;
    uint32_t* d_slot_list_ = nullptr;           // slot index of the i-th distinct junction
    TableRef table_ref() const { return TableRef{d_table_, table_slots_ - 1, d_slot_list_, table_slots_}; }
    Slot* d_spill_ = nullptr; uint32_t spill_cap_ = 0;
    Cand* d_cands_ = nullptr; uint32_t cand_cap_ = 0;
    uint64_t unique_upper_ = 0;                 // host-side upper bound of occupied slots
    uint64_t next_ord_ = 0;                     // one ordinal space per handle (first-seen names, last-writer strand): the next unused
                                                // alignment ordinal, shared by rtjx_add, rtjx_scan_batch and the BAM runs
    uint64_t run_ord_base_ = 0;                 // next_ord_ when the current BAM run started
    bool dirty_ = false;                        // device table changed since the last finalize

    // intron-motif mode: genome in HBM (one byte per base) and the BAM-tid -> sequence map
    bool genome_loaded_ = false;
    FastaGenome genome_;                        // names / offsets / lengths (bases are dropped after the upload)
    uint8_t* d_genome_ = nullptr;
    unsigned long long* d_g_off_ = nullptr; unsigned long long* d_g_len_ = nullptr; uint32_t g_n_ = 0, g_cap_ = 0;
    std::vector<std::string> genome_map_for_;   // contig list the device map was built for

    // batched variant regions: device copy sorted by (tid, beg) + per-region results in the caller's order
    int32_t* d_vr_tid_ = nullptr; int32_t* d_vr_beg_ = nullptr; int32_t* d_vr_end_ = nullptr; uint32_t vr_cap_ = 0;
    VariantRegions vr_{nullptr, nullptr, nullptr, 0, 0, 0};
    std::vector<uint32_t> vr_orig_;             // sorted index -> caller's index
    std::vector<std::vector<rtjx_junction>> region_tables_;
    std::vector<rtjx_junction> unique_;          // set<Junction> of the second caller, in its order (contig name, start, end)
    std::vector<uint32_t> unique_first_;         // region (variant) whose insert won
    std::vector<std::vector<uint32_t>> unique_regions_;   // junction_to_variant_: regions whose window holds the junction, ascending
    OutJunctionR* d_out_r_ = nullptr; uint32_t fin_r_cap_ = 0; void* d_ws_r_ = nullptr; size_t ws_r_cap_ = 0;
    OutJunctionR* h_final_r_ = nullptr; uint32_t h_final_r_cap_ = 0;
    int finalize_regions();
    int ensure_region_buffers(uint32_t n);

    // `-b` mode: the device table is keyed (junction, barcode id + 1); finalize folds it into a junction-level table and
    // keeps the pairs, sorted by (junction, first ordinal), in h_final_r_ for write_barcodes
    bool bc_mode_ = false;
    std::string bc_tag_ = "CB";
    BarcodeDict bc_dict_;
    Slot* d_fold_table_ = nullptr; uint32_t* d_fold_list_ = nullptr; uint32_t fold_slots_ = 0;
    uint32_t* d_fold_counters_ = nullptr;
    uint32_t bc_pairs_n_ = 0;
    int fold_barcode_table(uint32_t n_pairs, cudaStream_t st, TableRef* folded, uint32_t* n_junctions);

    // device batch ring for host-resident input
    struct DevBatch {
        int32_t* tid = nullptr; int32_t* pos = nullptr; uint32_t* meta = nullptr; uint32_t* cig_off = nullptr;
        uint32_t* cigar = nullptr; uint32_t* bc = nullptr; uint32_t cap_reads = 0, cap_ops = 0; cudaEvent_t free_ev = nullptr;
    };
    DevBatch dev_batch_[2];
    int dev_batch_next_ = 0;
    int ensure_dev_batch(DevBatch& d, uint32_t reads, uint32_t ops);

    // finalize scratch (grow-only; cudaMalloc/cudaFree per call cost ~20 ms on a loaded context)
    OutJunction* d_out_ = nullptr; uint32_t fin_cap_ = 0;
    void* d_ws_ = nullptr; size_t ws_cap_ = 0;
    uint32_t* d_rank_ = nullptr; size_t rank_cap_ = 0; bool rank_dirty_ = true;
    rtjx_junction* h_final_ = nullptr; uint32_t h_final_cap_ = 0;   // pinned D2H staging
    int ensure_finalize_buffers(uint32_t n, size_t n_contigs);

    // finalized table (host): in the pinned D2H buffer when it came straight from the device (no copy), else in final_
    std::vector<rtjx_junction> final_;
    uint32_t pinned_final_n_ = 0;
    const rtjx_junction* final_data() const { return pinned_final_n_ ? h_final_ : final_.data(); }
    size_t final_size() const { return pinned_final_n_ ? pinned_final_n_ : final_.size(); }
    std::vector<rtjx_junction> imported_;
    std::vector<size_t> import_sizes_;          // one entry per rtjx_import call (= one shard's table)
    bool merge_sorted_shards();
    bool finalized_ = false;

    // cached feeder output for load_batch
    struct LoadedBatch;
    std::unique_ptr<LoadedBatch> loaded_;

    // device-side feeder (BGZF inflate + record split on the GPU), device_run.cc
    struct DeviceFeed;
    struct DeviceFeedDeleter { void operator()(DeviceFeed* p) const; };
    std::unique_ptr<DeviceFeed, DeviceFeedDeleter> dfeed_;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> feed_prof_;
    int feed_seed_mode_ = 0;                    // record starts of a device run: 0 found on the device, 1 index (linear + bin chunks), 2 linear index only
    uint32_t feed_decline_flags_ = 0;           // FEED_FLAG_* of the group that made the device path decline

    // profiling
    struct ProfEv { cudaEvent_t a, b, c; };
    std::vector<ProfEv> prof_pending_;
    std::vector<cudaEvent_t> ev_pool_;
    cudaEvent_t get_event();

    rtjx_stats stats_;
};

}  // namespace rtjx
