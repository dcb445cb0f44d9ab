// regtools_b200/csrc/engine.cc — see engine.h.
#include "engine.h"
#include "buffer_cache.h"

#include <unistd.h>

#include <algorithm>
#include <thread>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <map>
#include <numeric>
#include <unordered_map>

namespace rtjx {

static_assert(sizeof(rtjx_junction) == sizeof(OutJunction), "ABI junction layout must match the device layout");
static_assert(offsetof(rtjx_junction, first_ord) == offsetof(OutJunction, first_ord), "layout");
static_assert(offsetof(rtjx_junction, name_index) == offsetof(OutJunction, name_index), "layout");

namespace {
double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
std::vector<uint32_t> contig_ranks(const std::vector<std::string>& names);      // defined with finalize below
uint32_t next_pow2(uint64_t x) {
    uint64_t p = 1;
    while (p < x) p <<= 1;
    return (uint32_t)std::min<uint64_t>(p, 1ull << 31);
}
}  // namespace

#define CK(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(RTJX_E_CUDA, std::string(#call " failed: ") + cudaGetErrorString(e__));      \
    } while (0)

Engine::Engine(const rtjx_params& p) : prm_(p) {
    bam_path_ = p.bam ? p.bam : "";
    region_ = p.region ? p.region : ".";
    tag_ = (p.strand_tag && p.strand_tag[0]) ? p.strand_tag : "XS";
    if (tag_.size() < 2) tag_.push_back('\0');
    fasta_path_ = p.fasta ? p.fasta : "";
    bc_mode_ = p.barcode_out != nullptr;                  // output_barcodes_file_ != "NA" (junctions_extractor.cc:204,393)
    if (p.barcode_tag && p.barcode_tag[0]) bc_tag_ = p.barcode_tag;
    if (bc_tag_.size() < 2) bc_tag_.push_back('\0');
    prm_.barcode_tag = nullptr;
    prm_.bam = prm_.region = prm_.strand_tag = prm_.fasta = prm_.barcode_out = nullptr;
    memset(&stats_, 0, sizeof stats_);
}

Engine::~Engine() {
    if (dev_ready_) {
        cudaSetDevice(prm_.device);
        cudaDeviceSynchronize();
        dfeed_.reset();
        for (auto& pe : prof_pending_) { ev_pool_.push_back(pe.a); ev_pool_.push_back(pe.b); ev_pool_.push_back(pe.c); }
        for (auto e : ev_pool_) cudaEventDestroy(e);
        for (auto& d : dev_batch_) {
            cached_dev_free(d.tid); cached_dev_free(d.pos); cached_dev_free(d.meta); cached_dev_free(d.cig_off); cached_dev_free(d.cigar);
            cached_dev_free(d.bc);
            if (d.free_ev) cudaEventDestroy(d.free_ev);
        }
        cached_dev_free(d_genome_); cached_dev_free(d_g_off_); cached_dev_free(d_g_len_);
        cached_dev_free(d_vr_tid_); cached_dev_free(d_vr_beg_); cached_dev_free(d_vr_end_);
        cached_dev_free(d_out_r_); cached_dev_free(d_ws_r_); cached_host_free(h_final_r_);
        cached_dev_free(d_fold_table_); cached_dev_free(d_fold_list_); cached_dev_free(d_fold_counters_);
        cached_dev_free(d_counters_); cached_host_free(h_counters_);
        cached_dev_free(d_table_); cached_dev_free(d_spill_); cached_dev_free(d_cands_);
        cached_dev_free(d_out_); cached_dev_free(d_ws_); cached_dev_free(d_rank_); cached_host_free(h_final_); cached_dev_free(d_slot_list_);
        if (stream_) cudaStreamDestroy(stream_);
        if (copy_stream_) cudaStreamDestroy(copy_stream_);
    }
}

ScanParams Engine::scan_params() const {
    ScanParams s;
    s.strandness = prm_.strandness; s.min_anchor = prm_.min_anchor;
    s.min_intron = prm_.min_intron; s.max_intron = prm_.max_intron;
    static const int env_variant = getenv("RTJX_SCAN_VARIANT") ? atoi(getenv("RTJX_SCAN_VARIANT")) : 0;
    static const int env_cfg = getenv("RTJX_SCAN_CFG") ? atoi(getenv("RTJX_SCAN_CFG")) : 0;
    s.variant = prm_.scan_variant ? prm_.scan_variant : (env_variant ? env_variant : 5);
    if (s.variant != 8) s.variant = 5;
    s.cfg = prm_.scan_cfg ? prm_.scan_cfg : env_cfg;
    s.genome = d_genome_; s.g_off = d_g_off_; s.g_len = d_g_len_; s.g_n = d_genome_ ? g_n_ : 0u;
    s.vr = vr_;
    return s;
}

cudaEvent_t Engine::get_event() {
    if (!ev_pool_.empty()) { cudaEvent_t e = ev_pool_.back(); ev_pool_.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

// There is no CPU fallback: a handle that has to compute needs a CUDA device.
int Engine::ensure_device() {
    if (dev_ready_) { cudaSetDevice(prm_.device); return RTJX_OK; }
    if (host_only()) return fail(RTJX_E_CUDA, "this handle was created host-only (device = -1); compute entry points need a CUDA device");
    const double t_dev0 = now_s();
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(RTJX_E_CUDA, std::string("no CUDA device available (") + cudaGetErrorString(e) +
                                     "); regtools-b200 has no CPU fallback for the junction kernels");
    if (prm_.device >= n) return fail(RTJX_E_CUDA, "CUDA device ordinal out of range");
    CK(cudaSetDevice(prm_.device));
    {
        int prio_lo = 0, prio_hi = 0;                    // (numerically lower = higher priority)
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaStreamCreateWithPriority(&stream_, cudaStreamNonBlocking, prio_hi));
    }
    CK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    CK(cached_dev_malloc(&d_counters_, CTR_COUNT * sizeof(uint32_t)));
    CK(cudaMemset(d_counters_, 0, CTR_COUNT * sizeof(uint32_t)));
    CK(cached_host_alloc(&h_counters_, 2 * CTR_COUNT * sizeof(uint32_t)));     // second half: counters of the `-b` fold table
    spill_cap_ = 4096;
    CK(cached_dev_malloc(&d_spill_, spill_cap_ * sizeof(Slot)));
    dev_ready_ = true;
    if (getenv("RTJX_TRACE")) fprintf(stderr, "[rtjx] device ready (CUDA context, streams, first allocations): %.1f ms\n", 1e3 * (now_s() - t_dev0));
    return RTJX_OK;
}

int Engine::sync_counters(cudaStream_t stream) {
    CK(cudaMemcpyAsync(h_counters_, d_counters_, CTR_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    stats_.d2h_bytes += CTR_COUNT * sizeof(uint32_t);
    if (h_counters_[CTR_CAND_OVERFLOW] || h_counters_[CTR_NSPILL])
        return fail(RTJX_E_STATE, "internal: candidate buffer or junction table overflowed");
    if (h_counters_[CTR_GENOME_MISS]) {               // get_reference_sequence threw (junctions_extractor.cc:553-555)
        const uint32_t s0 = h_counters_[CTR_GENOME_MISS_POS];
        return fail(RTJX_E_IO, std::string("Unable to extract FASTA sequence for position ") + contig((int32_t)h_counters_[CTR_GENOME_MISS] - 1) +
                                   ":" + std::to_string(s0 + 1u) + "-" + std::to_string(s0 + 2u) + "\n\n");
    }
    unique_upper_ = h_counters_[CTR_NUNIQUE];
    stats_.candidates = (uint64_t)h_counters_[CTR_TOTAL_CAND64 + 1] << 32 | h_counters_[CTR_TOTAL_CAND64];
    return RTJX_OK;
}

int Engine::ensure_cands(uint32_t n) {
    if (n <= cand_cap_) return RTJX_OK;
    CK(cudaDeviceSynchronize());
    cached_dev_free(d_cands_); d_cands_ = nullptr;
    uint32_t cap = std::max<uint32_t>(n + n / 4, 1u << 16);
    CK(cached_dev_malloc(&d_cands_, (size_t)cap * sizeof(Cand)));
    cand_cap_ = cap;
    return RTJX_OK;
}

// Keeps the open-addressed table at load <= 0.5 even if every incoming candidate were a new key,
// so an upsert can never run out of slots inside a kernel.
int Engine::ensure_table(uint32_t incoming, cudaStream_t stream) {
    if (!d_table_) {
        uint32_t want = 1u << (prm_.table_log2 ? std::min<uint32_t>(prm_.table_log2, 30) : 22);
        want = std::max(want, next_pow2(4ull * incoming));
        CK(cached_dev_malloc(&d_table_, (size_t)want * sizeof(Slot)));
        CK(cached_dev_malloc(&d_slot_list_, (size_t)want * sizeof(uint32_t)));
        CK(cudaMemsetAsync(d_table_, 0, (size_t)want * sizeof(Slot), stream));
        table_slots_ = want; unique_upper_ = 0;
    }
    if (2ull * (unique_upper_ + incoming) > table_slots_) {
        int rc = sync_counters(stream);               // tighten the bound with the real count
        if (rc) return rc;
        if (2ull * (unique_upper_ + incoming) > table_slots_) {
            uint32_t want = next_pow2(4ull * (unique_upper_ + incoming));
            Slot* nt = nullptr; uint32_t* nl = nullptr;
            CK(cached_dev_malloc(&nt, (size_t)want * sizeof(Slot)));
            CK(cached_dev_malloc(&nl, (size_t)want * sizeof(uint32_t)));
            CK(cudaMemsetAsync(nt, 0, (size_t)want * sizeof(Slot), stream));
            CK(cudaMemsetAsync(d_counters_ + CTR_NUNIQUE, 0, sizeof(uint32_t), stream));
            launch_table_rehash(d_table_, table_slots_, TableRef{nt, want - 1, nl, want}, d_counters_, stream);
            stats_.kernel_launches++;
            CK(cudaStreamSynchronize(stream));
            cached_dev_free(d_table_); cached_dev_free(d_slot_list_);
            d_table_ = nt; d_slot_list_ = nl; table_slots_ = want; stats_.table_grows++;
        }
    }
    unique_upper_ += incoming;
    return RTJX_OK;
}

// The genome goes to HBM once per handle; the tid -> sequence map follows the contig list (header of the BAM, or
// the names interned by add-only handles).
int Engine::ensure_genome() {
    if (fasta_path_.empty()) return RTJX_OK;
    if (!genome_loaded_) {
        std::string err;
        if (!load_fasta(fasta_path_, &genome_, &err)) return fail(RTJX_E_IO, err);
        CK(cached_dev_malloc(&d_genome_, genome_.bases.size()));
        CK(cudaMemcpy(d_genome_, genome_.bases.data(), genome_.bases.size(), cudaMemcpyHostToDevice));
        stats_.h2d_bytes += genome_.bases.size();
        std::vector<uint8_t>().swap(genome_.bases);
        genome_loaded_ = true;
    }
    if (genome_map_for_ != contigs_ || !d_g_off_) {
        const size_t n = std::max<size_t>(contigs_.size(), 1);
        if (n > g_cap_) {
            CK(cudaDeviceSynchronize());
            cached_dev_free(d_g_off_); cached_dev_free(d_g_len_); d_g_off_ = d_g_len_ = nullptr;
            g_cap_ = (uint32_t)(n * 2 + 64);
            CK(cached_dev_malloc(&d_g_off_, (size_t)g_cap_ * 8)); CK(cached_dev_malloc(&d_g_len_, (size_t)g_cap_ * 8));
        }
        std::vector<unsigned long long> off(n, 0ull), len(n, ~0ull);
        for (size_t t = 0; t < contigs_.size(); ++t) {
            const int q = genome_.find(contigs_[t]);
            if (q >= 0) { off[t] = genome_.offset[(size_t)q]; len[t] = genome_.length[(size_t)q]; }
        }
        CK(cudaDeviceSynchronize());                  // kernels in flight may still read the old map
        CK(cudaMemcpy(d_g_off_, off.data(), n * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_g_len_, len.data(), n * 8, cudaMemcpyHostToDevice));
        g_n_ = (uint32_t)contigs_.size();
        genome_map_for_ = contigs_;
    }
    return RTJX_OK;
}

// cand_bound = number of N ops in the batch when the producer knows it (the feeder counts them while
// copying CIGARs); 0 = unknown: size the candidate buffer for the worst case and read the real
// count back after cigar_scan (one stream synchronisation) before sizing the merge.
int Engine::process_device_batch(const BatchView& v, uint32_t cand_bound, cudaStream_t stream) {
    NvtxRange nvtx("rtjx:cigar_scan + junction_merge (enqueue)");
    int rc;
    const bool known = cand_bound != 0;
    next_ord_ = std::max<uint64_t>(next_ord_, v.first_ordinal + v.n_reads);
    if ((rc = ensure_genome())) return rc;
    const ScanParams sp = scan_params();
    if (sp.vr.n && ((reinterpret_cast<uintptr_t>(v.tid) | reinterpret_cast<uintptr_t>(v.pos) | reinterpret_cast<uintptr_t>(v.meta) |
                     reinterpret_cast<uintptr_t>(v.cig_off) | reinterpret_cast<uintptr_t>(v.cigar)) & 15u))
        return fail(RTJX_E_STATE, "internal: variant-region batches must be 16-byte aligned");
    if (sp.genome && ((reinterpret_cast<uintptr_t>(v.tid) | reinterpret_cast<uintptr_t>(v.pos) | reinterpret_cast<uintptr_t>(v.meta) |
                       reinterpret_cast<uintptr_t>(v.cig_off) | reinterpret_cast<uintptr_t>(v.cigar)) & 15u))
        return fail(RTJX_E_ARG, "device batch arrays must be 16-byte aligned when a FASTA is given");
    if (v.bc && ((reinterpret_cast<uintptr_t>(v.tid) | reinterpret_cast<uintptr_t>(v.pos) | reinterpret_cast<uintptr_t>(v.meta) |
                  reinterpret_cast<uintptr_t>(v.cig_off) | reinterpret_cast<uintptr_t>(v.cigar)) & 15u))
        return fail(RTJX_E_STATE, "internal: -b batches must be 16-byte aligned (only the tiled scan kernel knows barcodes)");
    if (sp.vr.n) {
        // variant-region mode: an alignment yields one candidate per region it belongs to, so the candidate count is only
        // known after the scan: scan, read the count back, grow and re-scan if the buffer was too small, then merge
        if ((rc = ensure_cands(std::max(std::max(cand_bound, v.n_ops / 4u) * 2u, 1u << 16) + cigar_scan_cand_slack()))) return rc;
        for (int attempt = 0;; ++attempt) {
            CK(cudaMemsetAsync(d_counters_ + CTR_NCAND, 0, 2 * sizeof(uint32_t), stream));       // NCAND + CAND_OVERFLOW
            launch_cigar_scan(v, sp, d_cands_, cand_cap_, d_counters_, stream);
            CK(cudaMemcpyAsync(h_counters_, d_counters_, CTR_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            stats_.kernel_launches++;
            if (h_counters_[CTR_NCAND] <= cand_cap_) break;
            if (attempt) return fail(RTJX_E_STATE, "internal: candidate buffer overflowed twice in variant-region mode");
            const uint32_t need = h_counters_[CTR_NCAND];
            CK(cudaMemsetAsync(d_counters_ + CTR_CAND_OVERFLOW, 0, sizeof(uint32_t), stream));
            if ((rc = ensure_cands(need + cigar_scan_cand_slack()))) return rc;
        }
        const uint32_t n_cand = h_counters_[CTR_NCAND];
        unique_upper_ = h_counters_[CTR_NUNIQUE];
        if (n_cand) {
            if ((rc = ensure_table(n_cand, stream))) return rc;
            launch_junction_merge(d_cands_, d_counters_ + CTR_NCAND, n_cand, sp, table_ref(), d_spill_, spill_cap_, d_counters_, stream);
            stats_.kernel_launches++;
        }
        CK(cudaGetLastError());
        stats_.reads += v.n_reads; stats_.cigar_ops += v.n_ops; stats_.batches++;
        dirty_ = true; finalized_ = false;
        return RTJX_OK;
    }
    // cigar_scan reserves candidate slots in chunks: the list holds up to cigar_scan_cand_slack() padding entries (tid = -1)
    const uint32_t slack = cigar_scan_cand_slack();
    if ((rc = ensure_cands((known ? cand_bound : std::max(v.n_ops, 1u)) + slack))) return rc;
    if (known && (rc = ensure_table(cand_bound, stream))) return rc;
    CK(cudaMemsetAsync(d_counters_ + CTR_NCAND, 0, sizeof(uint32_t), stream));
    ProfEv pe{nullptr, nullptr, nullptr};
    if (prm_.profile) { pe.a = get_event(); pe.b = get_event(); pe.c = get_event(); cudaEventRecord(pe.a, stream); }
    launch_cigar_scan(v, sp, d_cands_, cand_cap_, d_counters_, stream);
    if (prm_.profile) cudaEventRecord(pe.b, stream);
    if (!known) {
        if ((rc = sync_counters(stream))) return rc;     // also tightens unique_upper_
        cand_bound = h_counters_[CTR_NCAND];
        if ((rc = ensure_table(std::max(cand_bound, 1u), stream))) return rc;
    }
    launch_junction_merge(d_cands_, d_counters_ + CTR_NCAND, known ? cand_bound + slack : cand_bound, sp, table_ref(),
                          d_spill_, spill_cap_, d_counters_, stream);
    if (prm_.profile) { cudaEventRecord(pe.c, stream); prof_pending_.push_back(pe); }
    CK(cudaGetLastError());
    stats_.kernel_launches += (v.n_reads ? 1 : 0) + (cand_bound ? 1 : 0);   // cigar_scan, junction_merge
    stats_.reads += v.n_reads; stats_.cigar_ops += v.n_ops; stats_.batches++;
    dirty_ = true; finalized_ = false;
    if (prof_pending_.size() > 4096) resolve_profile_events();
    return RTJX_OK;
}

void Engine::resolve_profile_events() {
    for (auto& pe : prof_pending_) {
        cudaEventSynchronize(pe.c);
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pe.a, pe.b) == cudaSuccess) stats_.scan_ms += ms;
        if (cudaEventElapsedTime(&ms, pe.b, pe.c) == cudaSuccess) stats_.merge_ms += ms;
        ev_pool_.push_back(pe.a); ev_pool_.push_back(pe.b); ev_pool_.push_back(pe.c);
    }
    prof_pending_.clear();
}

int Engine::ensure_dev_batch(DevBatch& d, uint32_t reads, uint32_t ops) {
    if (!d.free_ev) CK(cudaEventCreateWithFlags(&d.free_ev, cudaEventDisableTiming));
    if (reads > d.cap_reads) {
        CK(cudaEventSynchronize(d.free_ev));
        cached_dev_free(d.tid); cached_dev_free(d.pos); cached_dev_free(d.meta); cached_dev_free(d.cig_off);
        uint32_t cap = std::max(reads, 1u << 12);
        CK(cached_dev_malloc(&d.tid, (size_t)cap * 4)); CK(cached_dev_malloc(&d.pos, (size_t)cap * 4));
        CK(cached_dev_malloc(&d.meta, (size_t)cap * 4)); CK(cached_dev_malloc(&d.cig_off, ((size_t)cap + 4) * 4));
        if (bc_mode_) { cached_dev_free(d.bc); d.bc = nullptr; CK(cached_dev_malloc(&d.bc, (size_t)cap * 4)); }
        d.cap_reads = cap;
    }
    if (ops > d.cap_ops) {
        CK(cudaEventSynchronize(d.free_ev));
        cached_dev_free(d.cigar);
        uint32_t cap = std::max(ops, 1u << 12);
        CK(cached_dev_malloc(&d.cigar, ((size_t)cap + 4) * 4));
        d.cap_ops = cap;
    }
    return RTJX_OK;
}

int Engine::scan_batch(const rtjx_batch& b, int location, cudaStream_t user_stream) {
    int rc = ensure_device();
    if (rc) return rc;
    if (bc_mode_ && !b.bc) return fail(RTJX_E_ARG, "a -b handle needs rtjx_batch.bc (ids from rtjx_intern_barcode)");
    if (b.n_reads == 0) return RTJX_OK;
    if (!b.tid || !b.pos || !b.meta || !b.cig_off || (b.n_ops && !b.cigar)) return fail(RTJX_E_ARG, "null array in batch");
    cudaStream_t st = user_stream ? user_stream : stream_;
    BatchView v;
    v.n_reads = b.n_reads; v.n_ops = b.n_ops; v.first_ordinal = b.first_ordinal;
    if (location == RTJX_LOC_DEVICE) {
        if ((reinterpret_cast<uintptr_t>(b.cigar) & 15u) != 0) return fail(RTJX_E_ARG, "device cigar array must be 16-byte aligned");
        v.tid = b.tid; v.pos = b.pos; v.meta = b.meta; v.cig_off = b.cig_off; v.cigar = b.cigar;
        if (bc_mode_) v.bc = b.bc;
    } else if (location == RTJX_LOC_HOST) {
        DevBatch& d = dev_batch_[dev_batch_next_]; dev_batch_next_ ^= 1;
        if ((rc = ensure_dev_batch(d, b.n_reads, b.n_ops))) return rc;
        CK(cudaStreamWaitEvent(st, d.free_ev, 0));
        CK(cudaMemcpyAsync(d.tid, b.tid, (size_t)b.n_reads * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d.pos, b.pos, (size_t)b.n_reads * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d.meta, b.meta, (size_t)b.n_reads * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d.cig_off, b.cig_off, ((size_t)b.n_reads + 1) * 4, cudaMemcpyHostToDevice, st));
        if (b.n_ops) CK(cudaMemcpyAsync(d.cigar, b.cigar, (size_t)b.n_ops * 4, cudaMemcpyHostToDevice, st));
        stats_.h2d_bytes += (size_t)b.n_reads * 16 + 4 + (size_t)b.n_ops * 4;
        v.tid = d.tid; v.pos = d.pos; v.meta = d.meta; v.cig_off = d.cig_off; v.cigar = d.cigar;
        if (bc_mode_) {
            CK(cudaMemcpyAsync(d.bc, b.bc, (size_t)b.n_reads * 4, cudaMemcpyHostToDevice, st));
            stats_.h2d_bytes += (size_t)b.n_reads * 4;
            v.bc = d.bc;
        }
        rc = process_device_batch(v, b.n_junction_ops, st);
        cudaEventRecord(d.free_ev, st);
        return rc;
    } else {
        return fail(RTJX_E_ARG, "location must be RTJX_LOC_HOST or RTJX_LOC_DEVICE");
    }
    return process_device_batch(v, b.n_junction_ops, st);
}

int Engine::add(const rtjx_candidate* c, size_t n) {
    int rc = ensure_device();
    if (rc) return rc;
    if (bc_mode_) return fail(RTJX_E_UNSUPPORTED, "rtjx_add carries no barcodes: a -b handle is fed by rtjx_run");
    if (n == 0) return RTJX_OK;
    if (!c) return fail(RTJX_E_ARG, "null candidates");
    // add_junction is called once per candidate in order; order is carried by the ordinal.
    const size_t CH = 1u << 20;
    std::vector<Cand> h;
    for (size_t o = 0; o < n; o += CH) {
        size_t m = std::min(CH, n - o);
        h.resize(m);
        for (size_t i = 0; i < m; ++i) {
            const rtjx_candidate& s = c[o + i];
            Cand& d = h[i];
            d.start = s.start; d.end = s.end; d.ts = s.thick_start; d.te = s.thick_end;
            d.ord = (next_ord_ + i) << 16; d.tid = s.tid; d.strand = s.strand;
        }
        if ((rc = ensure_cands((uint32_t)m))) return rc;
        if ((rc = ensure_table((uint32_t)m, stream_))) return rc;
        CK(cudaMemcpyAsync(d_cands_, h.data(), m * sizeof(Cand), cudaMemcpyHostToDevice, stream_));
        launch_junction_merge(d_cands_, nullptr, (uint32_t)m, scan_params(), table_ref(), d_spill_, spill_cap_, d_counters_, stream_);
        CK(cudaStreamSynchronize(stream_));
        next_ord_ += m; stats_.kernel_launches++; stats_.h2d_bytes += m * sizeof(Cand);
    }
    dirty_ = true; finalized_ = false;
    return RTJX_OK;
}

// ---- whole-file / region run ---------------------------------------------------------------------
struct EngineSink : BatchSink {
    static constexpr int NB = 3;
    Engine* e; HostBatch hb[NB]; cudaEvent_t done[NB]; bool in_flight[NB]; int next = 0; int rc = 0;
    uint32_t cap_reads, cap_ops;
    EngineSink(Engine* eng, uint32_t reads, uint32_t ops) : e(eng), cap_reads(reads), cap_ops(ops) {
        for (int i = 0; i < NB; ++i) { done[i] = nullptr; in_flight[i] = false; }
    }
    int init() {
        for (int i = 0; i < NB; ++i) {
            HostBatch& b = hb[i];
            if (cached_host_alloc(&b.tid, (size_t)cap_reads * 4) != cudaSuccess ||
                cached_host_alloc(&b.pos, (size_t)cap_reads * 4) != cudaSuccess ||
                cached_host_alloc(&b.meta, (size_t)cap_reads * 4) != cudaSuccess ||
                cached_host_alloc(&b.cig_off, ((size_t)cap_reads + 1) * 4) != cudaSuccess ||
                cached_host_alloc(&b.cigar, (size_t)cap_ops * 4) != cudaSuccess ||
                (e->bc_mode_ && cached_host_alloc(&b.bc, (size_t)cap_reads * 4) != cudaSuccess) ||
                cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess)
                return e->fail(RTJX_E_CUDA, "pinned batch allocation failed");
            b.cap_reads = cap_reads; b.cap_ops = cap_ops;
        }
        return RTJX_OK;
    }
    ~EngineSink() override {
        for (int i = 0; i < NB; ++i) {
            if (in_flight[i]) cudaEventSynchronize(done[i]);
            cached_host_free(hb[i].tid); cached_host_free(hb[i].pos); cached_host_free(hb[i].meta);
            cached_host_free(hb[i].cig_off); cached_host_free(hb[i].cigar);
            if (hb[i].bc) cached_host_free(hb[i].bc);
            if (done[i]) cudaEventDestroy(done[i]);
        }
    }
    HostBatch* acquire() override {
        int j = next; next = (next + 1) % NB;
        if (in_flight[j]) { double t0 = now_s(); cudaEventSynchronize(done[j]); in_flight[j] = false; e->stats_.host_wait_s += now_s() - t0; }
        return &hb[j];
    }
    void submit(HostBatch* b) override {
        if (rc || b->n_reads == 0) return;
        int j = (int)(b - hb);
        rc = push(*b, j);
    }
    int push(HostBatch& b, int j) {
        Engine::DevBatch& d = e->dev_batch_[e->dev_batch_next_]; e->dev_batch_next_ ^= 1;
        int r = e->ensure_dev_batch(d, b.n_reads, b.n_ops);
        if (r) return r;
        cudaStream_t cs = e->copy_stream_, ks = e->stream_;
        cudaStreamWaitEvent(cs, d.free_ev, 0);
        cudaMemcpyAsync(d.tid, b.tid, (size_t)b.n_reads * 4, cudaMemcpyHostToDevice, cs);
        cudaMemcpyAsync(d.pos, b.pos, (size_t)b.n_reads * 4, cudaMemcpyHostToDevice, cs);
        cudaMemcpyAsync(d.meta, b.meta, (size_t)b.n_reads * 4, cudaMemcpyHostToDevice, cs);
        cudaMemcpyAsync(d.cig_off, b.cig_off, ((size_t)b.n_reads + 1) * 4, cudaMemcpyHostToDevice, cs);
        if (b.n_ops) cudaMemcpyAsync(d.cigar, b.cigar, (size_t)b.n_ops * 4, cudaMemcpyHostToDevice, cs);
        if (e->bc_mode_) {
            if (!b.bc || !d.bc) return e->fail(RTJX_E_STATE, "internal: barcode column missing");
            if (e->bc_dict_.names.size() >= (1u << 24) - 1u)     // the id travels in 24 bits of the candidate
                return e->fail(RTJX_E_UNSUPPORTED, "more than 16.7 million distinct barcodes");
            cudaMemcpyAsync(d.bc, b.bc, (size_t)b.n_reads * 4, cudaMemcpyHostToDevice, cs);
            e->stats_.h2d_bytes += (size_t)b.n_reads * 4;
        }
        cudaEventRecord(done[j], cs);
        in_flight[j] = true;
        e->stats_.h2d_bytes += (size_t)b.n_reads * 16 + 4 + (size_t)b.n_ops * 4;
        cudaStreamWaitEvent(ks, done[j], 0);
        BatchView v;
        v.n_reads = b.n_reads; v.n_ops = b.n_ops; v.first_ordinal = e->run_ord_base_ + b.first_ordinal;
        v.tid = d.tid; v.pos = d.pos; v.meta = d.meta; v.cig_off = d.cig_off; v.cigar = d.cigar;
        v.bc = e->bc_mode_ ? d.bc : nullptr;
        r = e->process_device_batch(v, b.n_junction_ops, ks);
        cudaEventRecord(d.free_ev, ks);
        return r;
    }
};

int Engine::open_bam(std::unique_ptr<BamFile>* bam, BaiIndex* idx, IterSpec* spec) {
    if (bam_path_.empty()) return fail(RTJX_E_ARG, "no BAM given");
    std::string err;
    bam->reset(new BamFile());
    if (access(bam_path_.c_str(), R_OK) != 0) return fail(RTJX_E_OPEN_BAM, "Unable to open BAM/SAM file.\n\n");
    bool csi = false;
    const bool have_idx = BaiIndex::load_for_bam(bam_path_, idx, &csi);
    if (!(*bam)->open(bam_path_, &err)) {
        // the reference opens any readable file and fails later at the index or the header
        if (!have_idx && !csi) return fail(RTJX_E_OPEN_INDEX, "Unable to open BAM/SAM index. Make sure alignments are indexed\n\n");
        return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
    }
    if (!have_idx) {
        return fail(RTJX_E_OPEN_INDEX, "Unable to open BAM/SAM index. Make sure alignments are indexed\n\n");
    }
    contigs_ = (*bam)->header().names; rank_dirty_ = true;
    if (!parse_region(**bam, region_, spec)) return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
    if (prm_.shard_world > 1) {
        if (spec->kind != IterSpec::WholeFile) return fail(RTJX_E_ARG, "contig sharding needs region \".\"");
        std::vector<int32_t> assign = plan_contig_shards(**bam, *idx, prm_.shard_world);
        spec->kind = IterSpec::Contigs;
        for (size_t t = 0; t < assign.size(); ++t) if (assign[t] == prm_.shard_rank) spec->contigs.push_back((int32_t)t);
    } else if (spec->kind == IterSpec::WholeFile) {
        uint64_t off0;
        if (!idx->whole_file_start(&off0)) return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
    } else if (spec->kind == IterSpec::Region) {
        if (spec->end < spec->beg || (size_t)spec->tid >= idx->refs.size())
            return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
    }
    return RTJX_OK;
}

int Engine::run() {
    int rc = run_impl();
    // with a FASTA the reference throws inside identify_junctions_from_BAM when a junction lies on a contig the FASTA lacks
    if (rc == RTJX_OK && !fasta_path_.empty() && dev_ready_) rc = sync_counters(stream_);
    return rc;
}

int Engine::run_impl() {
    NvtxRange nvtx("rtjx:run (identify_junctions_from_BAM)");
    const double t_start = now_s();
    run_ord_base_ = next_ord_;                  // alignments of this run are numbered after everything the handle has seen (rtjx_add included)
    std::unique_ptr<BamFile> bam; BaiIndex idx; IterSpec spec;
    int rc = open_bam(&bam, &idx, &spec);
    if (rc) return rc;
    if ((rc = ensure_device())) return rc;
    // Whole-file and contig-shard runs inflate and split records on the GPU (inflate_mode 0 = auto, 2 = force);
    // regions, tiny files and anything the device path declines go through the host feeder below.
    bool streamable = spec.kind == IterSpec::WholeFile || spec.kind == IterSpec::Contigs;
    // A `-r` region that spans enough of the file goes the same way (round 2; it used to be the single-threaded, htslib-shaped
    // reader below whatever its size): the byte span from the first to the last index chunk of the query streams through the
    // device feeder, and cigar_scan keeps the alignments hts_itr_next would return — tid equal, pos < end, endpos > beg
    // (hts.c:1941-1963) — with the variant-region test of rtjx_run_regions on ONE region that tags nothing.  Equivalent for an
    // index that is consistent with its file: no alignment in front of the linear index's offset overlaps the region, every
    // overlapping alignment behind it lies in a chunk of the query, and in a sorted file nothing after the iterator's stop
    // (tid changes or pos >= end) passes the test.  Anything odd makes the device feeder decline, and the reader below decides.
    struct FilterGuard { Engine* e; bool on; ~FilterGuard() { if (on) e->vr_ = VariantRegions{nullptr, nullptr, nullptr, 0, 0, 0}; } } filter_guard{this, false};
    if (spec.kind == IterSpec::Region && !bc_mode_ && vr_.n == 0 && spec.end >= spec.beg && (size_t)spec.tid < idx.refs.size()) {
        const std::vector<Chunk64> off = idx.query(spec.tid, spec.beg, spec.end);
        static const uint64_t min_span = [] { const char* v = getenv("RTJX_REGION_DEVICE_MB"); return (uint64_t)(v ? atoi(v) : 8) << 20; }();
        if (!off.empty() && (prm_.inflate_mode == 2 || (prm_.inflate_mode == 0 && (off.back().end >> 16) - (off.front().beg >> 16) >= min_span))) {
            if (vr_cap_ == 0) {
                vr_cap_ = 64;
                CK(cached_dev_malloc(&d_vr_tid_, (size_t)vr_cap_ * 4)); CK(cached_dev_malloc(&d_vr_beg_, (size_t)vr_cap_ * 4));
                CK(cached_dev_malloc(&d_vr_end_, (size_t)vr_cap_ * 4));
            }
            const int32_t t = spec.tid, b = (int32_t)std::max<int64_t>(spec.beg, 0), e = (int32_t)std::min<int64_t>(spec.end, INT32_MAX);
            CK(cudaMemcpy(d_vr_tid_, &t, 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_vr_beg_, &b, 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_vr_end_, &e, 4, cudaMemcpyHostToDevice));
            vr_ = VariantRegions{d_vr_tid_, d_vr_beg_, d_vr_end_, 1u, (uint32_t)std::max<int64_t>((int64_t)e - b, 1), 0u};
            filter_guard.on = true;
            streamable = true;
        }
    }
    // `-b`: the barcode strings are dictionary-encoded by the host feeder, so that mode never takes the device feeder
    if (bc_mode_ && prm_.shard_world > 1) return fail(RTJX_E_UNSUPPORTED, "-b barcodes are not exchanged between contig shards");
    if (!bc_mode_ && streamable && (prm_.inflate_mode == 2 || (prm_.inflate_mode == 0 && bam->size() >= (1u << 20)))) {
        const rtjx_stats saved = stats_;
        for (int attempt = 0; attempt < 3; ++attempt) {
            // Record starts: first found on the device, one per BGZF block, verified by the chain walk; if that run is declined
            // once more with every record start the index knows (linear index + bin chunk boundaries); if that is declined too —
            // e.g. an index whose chunk ends are not record boundaries — with the linear index alone; then the host feeder.
            feed_seed_mode_ = attempt;
            feed_decline_flags_ = 0;
            rc = run_device(*bam, idx, spec);
            if (rc == RTJX_OK) return RTJX_OK;
            if (rc < 0) return rc;
            if (getenv("RTJX_TRACE")) fprintf(stderr, "[rtjx] device feed declined (seed mode %d, flags %u)\n", attempt, feed_decline_flags_);
            if ((rc = clear())) return rc;                    // declined: start over
            stats_ = saved;
            contigs_ = bam->header().names; rank_dirty_ = true;
            // (a walk started at a bogus seed can also end in a "malformed record" report, so any decline is retried; a
            // genuinely malformed file costs two more device passes before the host path below)
        }
    }
    if (filter_guard.on) { vr_ = VariantRegions{nullptr, nullptr, nullptr, 0, 0, 0}; filter_guard.on = false; }   // the reader below applies the iterator's own test
    uint32_t reads = prm_.batch_reads ? prm_.batch_reads : (spec.kind == IterSpec::Region ? (1u << 15) : (1u << 20));
    reads = std::max(reads, 1024u);
    uint32_t ops = std::max<uint32_t>(2 * reads, 1u << 17);      // one read may carry 65535 ops
    EngineSink sink(this, reads, ops);
    if ((rc = sink.init())) return rc;
    FeederOptions fo;
    fo.n_threads = prm_.n_threads; fo.xs_mode = prm_.strandness == 0; fo.tag[0] = tag_[0]; fo.tag[1] = tag_[1];
    if (bc_mode_) { fo.barcodes = &bc_dict_; fo.bc_tag[0] = bc_tag_[0]; fo.bc_tag[1] = bc_tag_[1]; }
    FeederStats fs; std::string err;
    uint64_t reads_before = stats_.reads;
    if (!feed_alignments(*bam, idx, spec, fo, &sink, &fs, &err)) return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
    if (sink.rc) return sink.rc;
    CK(cudaStreamSynchronize(copy_stream_));
    CK(cudaStreamSynchronize(stream_));
    (void)reads_before;
    if (bc_mode_ && bc_dict_.bad_type)      // bam_aux2Z returns NULL (sam.c:1309-1315) and std::string(NULL) ends the reference
        return fail(RTJX_E_IO, "barcode tag " + bc_tag_.substr(0, 2) + " is present with a non-string type (the reference aborts on such a file)");
    stats_.bgzf_blocks += fs.bgzf_blocks; stats_.compressed_bytes += fs.compressed_bytes; stats_.inflated_bytes += fs.inflated_bytes;
    stats_.host_inflate_s += fs.inflate_s; stats_.host_parse_s += fs.parse_s; stats_.host_wait_s += fs.wait_s;
    stats_.total_s += now_s() - t_start;
    return RTJX_OK;
}

// ---- batched variant regions (second caller of the path) --------------------------------------------
// The reference builds one JunctionsExtractor per variant (cis_splice_effects_identifier.cc:288-290): open, index load,
// iterate the region, merge, sort — 50k times on the same BAM.  Here every region becomes a row of a sorted interval
// table in HBM, ONE pass over the file feeds cigar_scan, which emits a candidate once per region its alignment belongs
// to, and the region index is part of the junction key; finalize ranks the names per region (each region is its own
// extractor with its own JUNC numbering) and sorts per region.
int Engine::run_regions(const char* const* regions, size_t n) {
    if (n && !regions) return fail(RTJX_E_ARG, "null region list");
    if (n >= (1u << 29)) return fail(RTJX_E_ARG, "too many regions");
    if (region_ != ".") return fail(RTJX_E_ARG, "rtjx_run_regions needs a handle created with region \".\"");
    if (bc_mode_) return fail(RTJX_E_UNSUPPORTED, "rtjx_run_regions has no barcode mode (the 8-arg ctor never sets -b, junctions_extractor.h:199-205)");
    std::unique_ptr<BamFile> bam; BaiIndex idx; IterSpec spec;
    int rc = open_bam(&bam, &idx, &spec);
    if (rc) return rc;
    if ((rc = ensure_device())) return rc;
    struct R { int32_t tid, beg, end; uint32_t orig; };
    std::vector<R> rs(n);
    uint32_t max_len = 1;
    for (size_t i = 0; i < n; ++i) {
        IterSpec one;
        if (!regions[i] || !parse_region(*bam, regions[i], &one) || one.kind != IterSpec::Region || one.end < one.beg ||
            (size_t)one.tid >= idx.refs.size())
            return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
        const int64_t b = std::max<int64_t>(one.beg, 0), e = std::min<int64_t>(one.end, INT32_MAX);
        rs[i] = R{one.tid, (int32_t)b, (int32_t)e, (uint32_t)i};
        max_len = std::max<uint32_t>(max_len, (uint32_t)(e - b));
    }
    std::sort(rs.begin(), rs.end(), [](const R& a, const R& b) { return a.tid != b.tid ? a.tid < b.tid : (a.beg != b.beg ? a.beg < b.beg : a.orig < b.orig); });
    region_tables_.assign(n, {});
    vr_orig_.resize(n);
    if ((rc = clear())) return rc;
    if (n == 0) return RTJX_OK;
    if (n > vr_cap_) {
        CK(cudaDeviceSynchronize());
        cached_dev_free(d_vr_tid_); cached_dev_free(d_vr_beg_); cached_dev_free(d_vr_end_); d_vr_tid_ = d_vr_beg_ = d_vr_end_ = nullptr;
        vr_cap_ = (uint32_t)(n + n / 4 + 64);
        CK(cached_dev_malloc(&d_vr_tid_, (size_t)vr_cap_ * 4)); CK(cached_dev_malloc(&d_vr_beg_, (size_t)vr_cap_ * 4));
        CK(cached_dev_malloc(&d_vr_end_, (size_t)vr_cap_ * 4));
    }
    {
        std::vector<int32_t> t(n), b(n), e(n);
        for (size_t i = 0; i < n; ++i) { t[i] = rs[i].tid; b[i] = rs[i].beg; e[i] = rs[i].end; vr_orig_[i] = rs[i].orig; }
        CK(cudaMemcpy(d_vr_tid_, t.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_vr_beg_, b.data(), n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d_vr_end_, e.data(), n * 4, cudaMemcpyHostToDevice));
        stats_.h2d_bytes += n * 12;
    }
    vr_ = VariantRegions{d_vr_tid_, d_vr_beg_, d_vr_end_, (uint32_t)n, max_len, 1u};
    rc = run();                                        // whole-file pass; the scan kernel does the region membership
    if (rc == RTJX_OK) rc = finalize_regions();
    vr_ = VariantRegions{nullptr, nullptr, nullptr, 0, 0, 0};
    const int rc2 = clear();                           // the handle's own table stays empty: results live in region_tables_
    return rc ? rc : rc2;
}

int Engine::ensure_region_buffers(uint32_t n) {
    if (n > fin_r_cap_) {
        CK(cudaDeviceSynchronize());
        cached_dev_free(d_out_r_); cached_dev_free(d_ws_r_); d_out_r_ = nullptr; d_ws_r_ = nullptr;
        const uint32_t cap = std::max<uint32_t>(n + n / 2, 1u << 14);
        CK(cached_dev_malloc(&d_out_r_, (size_t)cap * sizeof(OutJunctionR)));
        ws_r_cap_ = finalize_sort_regions_workspace_bytes(cap);
        CK(cached_dev_malloc(&d_ws_r_, ws_r_cap_));
        fin_r_cap_ = cap;
    }
    if (n > h_final_r_cap_) {
        cached_host_free(h_final_r_); h_final_r_ = nullptr;
        const uint32_t cap = std::max<uint32_t>(n + n / 2, 1u << 14);
        CK(cached_host_alloc(&h_final_r_, (size_t)cap * sizeof(OutJunctionR)));
        h_final_r_cap_ = cap;
    }
    return RTJX_OK;
}

int Engine::finalize_regions() {
    cudaStream_t st = stream_;
    int rc = sync_counters(st);
    if (rc) return rc;
    const uint32_t n = d_table_ ? h_counters_[CTR_NUNIQUE] : 0u;
    if (!n) return RTJX_OK;
    if ((rc = ensure_region_buffers(n))) return rc;
    if ((rc = ensure_finalize_buffers(1, contigs_.size()))) return rc;     // contig ranks
    if (rank_dirty_) {
        std::vector<uint32_t> cr = contig_ranks(contigs_);
        if (!cr.empty()) CK(cudaMemcpy(d_rank_, cr.data(), cr.size() * 4, cudaMemcpyHostToDevice));
        rank_dirty_ = false;
    }
    launch_table_compact_regions(table_ref(), n, d_out_r_, st);
    launch_finalize_sort_regions(d_out_r_, n, d_rank_, (uint32_t)contigs_.size(), d_ws_r_, ws_r_cap_, st);
    CK(cudaMemcpyAsync(h_final_r_, d_out_r_, (size_t)n * sizeof(OutJunctionR), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    stats_.kernel_launches += 2; stats_.d2h_bytes += (size_t)n * sizeof(OutJunctionR);
    for (uint32_t i = 0; i < n; ++i) {
        const OutJunctionR& r = h_final_r_[i];
        if (r.region == 0 || r.region > vr_orig_.size()) return fail(RTJX_E_STATE, "internal: junction without a region");
        rtjx_junction j;
        memcpy(&j, &r.j, sizeof j);
        region_tables_[vr_orig_[r.region - 1]].push_back(j);
    }
    return RTJX_OK;
}

int64_t Engine::region_count(size_t i) {
    if (i >= region_tables_.size()) return fail(RTJX_E_ARG, "region index out of range");
    return (int64_t)region_tables_[i].size();
}

int64_t Engine::region_get(size_t i, rtjx_junction* out, size_t cap) {
    if (i >= region_tables_.size()) return fail(RTJX_E_ARG, "region index out of range");
    const std::vector<rtjx_junction>& t = region_tables_[i];
    if (out && cap) memcpy(out, t.data(), std::min(cap, t.size()) * sizeof(rtjx_junction));
    return (int64_t)t.size();
}

// ---- the second caller's unique-junction set ---------------------------------------------------------
// cis_splice_effects_identifier.cc:288-299: for every variant, in order, the junctions of its region (get_all_junctions: sorted
// by compare_junctions) whose start OR end lies inside the variant's window [cis_effect_start, cis_effect_end] (both bounds
// inclusive, :294-295) are inserted into set<Junction> and map<Junction, set<variant>>.  Junction has no operator<: both
// containers compare through the implicit conversion to AnnotatedJunction (junctions_annotator.h:155-177) — contig NAME,
// start, end, strand-blind — and set::insert keeps the element that came first.
int Engine::unique_build(const uint32_t* win_start, const uint32_t* win_end, size_t n) {
    if (n != region_tables_.size()) return fail(RTJX_E_ARG, "one window per region of the last rtjx_run_regions call is needed");
    if (n && (!win_start || !win_end)) return fail(RTJX_E_ARG, "null window arrays");
    const std::vector<uint32_t> cr = contig_ranks(contigs_);
    struct Key {
        uint32_t crank, start, end;
        bool operator<(const Key& o) const { return crank != o.crank ? crank < o.crank : (start != o.start ? start < o.start : end < o.end); }
    };
    std::map<Key, size_t> index;                 // -> position in first-insert order
    std::vector<rtjx_junction> js; std::vector<uint32_t> first; std::vector<std::vector<uint32_t>> regs;
    for (size_t i = 0; i < n; ++i) {
        const uint32_t cs = win_start[i], ce = win_end[i];
        for (const rtjx_junction& j : region_tables_[i]) {
            if (!((j.start >= cs && j.start <= ce) || (j.end <= ce && j.end >= cs))) continue;
            const Key k{(uint32_t)j.tid < cr.size() ? cr[(size_t)j.tid] : 0x40000000u + (uint32_t)j.tid, j.start, j.end};
            auto it = index.find(k);
            if (it == index.end()) {
                index.emplace(k, js.size());
                js.push_back(j); first.push_back((uint32_t)i); regs.emplace_back(1, (uint32_t)i);
            } else if (regs[it->second].back() != (uint32_t)i) {
                regs[it->second].push_back((uint32_t)i);
            }
        }
    }
    unique_.clear(); unique_first_.clear(); unique_regions_.clear();
    unique_.reserve(js.size());
    for (const auto& kv : index) {               // set order
        unique_.push_back(js[kv.second]); unique_first_.push_back(first[kv.second]); unique_regions_.push_back(std::move(regs[kv.second]));
    }
    return RTJX_OK;
}
int64_t Engine::unique_count() { return (int64_t)unique_.size(); }
int64_t Engine::unique_get(rtjx_junction* out, uint32_t* first_region, size_t cap) {
    const size_t m = std::min(cap, unique_.size());
    if (out && m) memcpy(out, unique_.data(), m * sizeof(rtjx_junction));
    if (first_region && m) memcpy(first_region, unique_first_.data(), m * sizeof(uint32_t));
    return (int64_t)unique_.size();
}
int64_t Engine::unique_regions(size_t i, uint32_t* out, size_t cap) {
    if (i >= unique_regions_.size()) return fail(RTJX_E_ARG, "unique junction index out of range");
    const std::vector<uint32_t>& v = unique_regions_[i];
    if (out && cap) memcpy(out, v.data(), std::min(cap, v.size()) * sizeof(uint32_t));
    return (int64_t)v.size();
}

// ---- device inflate test hook ----------------------------------------------------------------------
int Engine::inflate_file(uint64_t max_blocks, void* out, uint64_t cap, uint64_t* out_len) {
    if (!out_len) return fail(RTJX_E_ARG, "out_len must not be NULL");
    int rc = ensure_device();
    if (rc) return rc;
    BamFile bam; std::string err;
    if (bam_path_.empty() || !bam.open(bam_path_, &err)) return fail(RTJX_E_OPEN_BAM, "Unable to open BAM/SAM file.\n\n");
    std::vector<BgzfBlockInfo> blocks; bool stop = false;
    scan_bgzf_blocks(bam, 0, bam.size(), max_blocks ? (size_t)max_blocks : (size_t)-1, UINT64_MAX, &blocks, &stop);
    std::vector<BgzfBlockDesc> desc(blocks.size());
    uint64_t in_total = 0, out_total = 0;
    for (size_t i = 0; i < blocks.size(); ++i) {
        desc[i].in_off = (uint32_t)in_total; desc[i].in_len = blocks[i].csize - 26;
        desc[i].out_off = (uint32_t)out_total; desc[i].out_len = blocks[i].isize;
        in_total += (desc[i].in_len + 3u) & ~3u; out_total += blocks[i].isize;
    }
    if (in_total > 0xfff00000ull || out_total > 0xfff00000ull) return fail(RTJX_E_ARG, "file too large for the single-shot inflate hook");
    *out_len = out_total;
    if (out_total > cap || !out) return blocks.empty() ? RTJX_OK : (out ? fail(RTJX_E_ARG, "output buffer too small") : RTJX_OK);
    std::vector<uint8_t> packed(in_total + 64, 0);
    for (size_t i = 0; i < blocks.size(); ++i) memcpy(packed.data() + desc[i].in_off, bam.data() + blocks[i].coff + 18, desc[i].in_len);
    uint8_t *d_in = nullptr, *d_out = nullptr; BgzfBlockDesc* d_desc = nullptr; uint32_t* d_status = nullptr;
    CK(cached_dev_malloc(&d_in, packed.size())); CK(cached_dev_malloc(&d_out, out_total + 16));
    CK(cached_dev_malloc(&d_desc, desc.size() * sizeof(BgzfBlockDesc) + 16)); CK(cached_dev_malloc(&d_status, desc.size() * 4 + 16));
    CK(cudaMemcpyAsync(d_in, packed.data(), packed.size(), cudaMemcpyHostToDevice, stream_));
    CK(cudaMemcpyAsync(d_desc, desc.data(), desc.size() * sizeof(BgzfBlockDesc), cudaMemcpyHostToDevice, stream_));
    cudaEvent_t ea = get_event(), eb = get_event();
    cudaEventRecord(ea, stream_);
    void* d_scratch = nullptr;
    CK(cached_dev_malloc(&d_scratch, bgzf_inflate_scratch_bytes((uint32_t)desc.size())));
    launch_bgzf_inflate(d_in, d_desc, (uint32_t)desc.size(), d_out, d_status, d_scratch, stream_);
    cudaEventRecord(eb, stream_);
    std::vector<uint32_t> status(desc.size());
    CK(cudaMemcpyAsync(status.data(), d_status, desc.size() * 4, cudaMemcpyDeviceToHost, stream_));
    CK(cudaMemcpyAsync(out, d_out, out_total, cudaMemcpyDeviceToHost, stream_));
    CK(cudaStreamSynchronize(stream_));
    CK(cudaGetLastError());
    float ms = 0; cudaEventElapsedTime(&ms, ea, eb); stats_.inflate_kernel_ms += ms;
    ev_pool_.push_back(ea); ev_pool_.push_back(eb);
    stats_.kernel_launches++; stats_.bgzf_blocks += blocks.size(); stats_.compressed_bytes += in_total; stats_.inflated_bytes += out_total;
    cached_dev_free(d_in); cached_dev_free(d_out); cached_dev_free(d_desc); cached_dev_free(d_status); cached_dev_free(d_scratch);
    for (size_t i = 0; i < status.size(); ++i)
        if (status[i]) return fail(RTJX_E_IO, "device inflate failed on BGZF block " + std::to_string(i) + " (code " + std::to_string(status[i]) + ")");
    return RTJX_OK;
}

// ---- feeder-only: SoA arrays for kernel-level tests and benches ----------------------------------
struct Engine::LoadedBatch {
    std::vector<int32_t> tid, pos; std::vector<uint32_t> meta, off, cigar;
};

namespace {
struct VectorSink : BatchSink {
    HostBatch hb; std::vector<int32_t> tid, pos; std::vector<uint32_t> meta, off, cig;
    std::vector<int32_t>* otid; std::vector<int32_t>* opos; std::vector<uint32_t>* ometa; std::vector<uint32_t>* ooff; std::vector<uint32_t>* ocig;
    VectorSink(uint32_t reads, uint32_t ops) : tid(reads), pos(reads), meta(reads), off(reads + 1), cig(ops) {
        hb.tid = tid.data(); hb.pos = pos.data(); hb.meta = meta.data(); hb.cig_off = off.data(); hb.cigar = cig.data();
        hb.cap_reads = reads; hb.cap_ops = ops;
    }
    std::vector<uint32_t> bcv; std::vector<uint32_t>* obc = nullptr;       // `-b` mode: per-alignment barcode ids
    void want_barcodes(std::vector<uint32_t>* out) { bcv.resize(hb.cap_reads); hb.bc = bcv.data(); obc = out; }
    HostBatch* acquire() override { return &hb; }
    void submit(HostBatch* b) override {
        uint32_t base = (uint32_t)ocig->size();
        if (obc) obc->insert(obc->end(), b->bc, b->bc + b->n_reads);
        otid->insert(otid->end(), b->tid, b->tid + b->n_reads);
        opos->insert(opos->end(), b->pos, b->pos + b->n_reads);
        ometa->insert(ometa->end(), b->meta, b->meta + b->n_reads);
        for (uint32_t i = 0; i < b->n_reads; ++i) ooff->push_back(base + b->cig_off[i]);
        ocig->insert(ocig->end(), b->cigar, b->cigar + b->n_ops);
    }
};
}  // namespace

int Engine::load_batch(uint64_t* n_reads, uint64_t* n_ops, int32_t* tid, int32_t* pos, uint32_t* meta,
                       uint32_t* cig_off, uint32_t* cigar) {
    if (!n_reads || !n_ops) return fail(RTJX_E_ARG, "n_reads/n_ops must not be NULL");
    if (!loaded_) {
        std::unique_ptr<BamFile> bam; BaiIndex idx; IterSpec spec;
        int rc = open_bam(&bam, &idx, &spec);
        if (rc) return rc;
        loaded_.reset(new LoadedBatch());
        VectorSink sink(1u << 18, 1u << 20);
        sink.otid = &loaded_->tid; sink.opos = &loaded_->pos; sink.ometa = &loaded_->meta; sink.ooff = &loaded_->off; sink.ocig = &loaded_->cigar;
        FeederOptions fo;
        fo.n_threads = prm_.n_threads; fo.xs_mode = prm_.strandness == 0; fo.tag[0] = tag_[0]; fo.tag[1] = tag_[1];
        FeederStats fs; std::string err;
        if (!feed_alignments(*bam, idx, spec, fo, &sink, &fs, &err)) { loaded_.reset(); return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n"); }
        if (loaded_->cigar.size() > 0xffffffffull || loaded_->tid.size() > 0xfffffffeull) { loaded_.reset(); return fail(RTJX_E_ARG, "too many alignments for one batch; use a region"); }
        loaded_->off.push_back((uint32_t)loaded_->cigar.size());
    }
    *n_reads = loaded_->tid.size(); *n_ops = loaded_->cigar.size();
    if (tid && pos && meta && cig_off && (cigar || loaded_->cigar.empty())) {
        memcpy(tid, loaded_->tid.data(), loaded_->tid.size() * 4);
        memcpy(pos, loaded_->pos.data(), loaded_->pos.size() * 4);
        memcpy(meta, loaded_->meta.data(), loaded_->meta.size() * 4);
        memcpy(cig_off, loaded_->off.data(), loaded_->off.size() * 4);
        if (!loaded_->cigar.empty()) memcpy(cigar, loaded_->cigar.data(), loaded_->cigar.size() * 4);
        loaded_.reset();
    }
    return RTJX_OK;
}

// Host feeder only (no device needed): the barcode dictionary ids of every alignment the handle's region iterates, as the
// device batches of a -b run carry them; the dictionary stays in the handle (barcode_name / barcode_stats).
int64_t Engine::load_barcodes(uint32_t* ids, size_t cap) {
    if (!bc_mode_) return fail(RTJX_E_STATE, "the handle was not created in -b mode (rtjx_params.barcode_out)");
    if (dirty_ || finalized_) return fail(RTJX_E_STATE, "rtjx_load_barcodes rebuilds the dictionary: use a fresh or cleared handle");
    std::unique_ptr<BamFile> bam; BaiIndex idx; IterSpec spec;
    int rc = open_bam(&bam, &idx, &spec);
    if (rc) return rc;
    LoadedBatch lb; std::vector<uint32_t> bc;
    VectorSink sink(1u << 16, 1u << 18);
    sink.otid = &lb.tid; sink.opos = &lb.pos; sink.ometa = &lb.meta; sink.ooff = &lb.off; sink.ocig = &lb.cigar;
    sink.want_barcodes(&bc);
    bc_dict_.clear();
    FeederOptions fo;
    fo.n_threads = prm_.n_threads; fo.xs_mode = prm_.strandness == 0; fo.tag[0] = tag_[0]; fo.tag[1] = tag_[1];
    fo.barcodes = &bc_dict_; fo.bc_tag[0] = bc_tag_[0]; fo.bc_tag[1] = bc_tag_[1];
    FeederStats fs; std::string err;
    if (!feed_alignments(*bam, idx, spec, fo, &sink, &fs, &err)) return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
    if (ids && cap) memcpy(ids, bc.data(), std::min(cap, bc.size()) * sizeof(uint32_t));
    return (int64_t)bc.size();
}

// ---- finalize ------------------------------------------------------------------------------------
namespace {
// rank of each contig name in std::string order, equal names get equal ranks (compare_junctions,
// junctions_extractor.h:120-126 compares the chrom strings)
std::vector<uint32_t> contig_ranks(const std::vector<std::string>& names) {
    std::vector<uint32_t> order(names.size()), rank(names.size());
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return names[a] < names[b]; });
    uint32_t r = 0;
    for (size_t i = 0; i < order.size(); ++i) {
        if (i && names[order[i]] != names[order[i - 1]]) ++r;
        rank[order[i]] = r;
    }
    return rank;
}
inline int name_cmp(uint32_t a, uint32_t b) {        // std::string compare of "JUNC%08d"
    if (a == b) return 0;
    if (a < 100000000u && b < 100000000u) return a < b ? -1 : 1;
    char sa[16], sb[16];
    snprintf(sa, sizeof sa, "%08d", (int)a); snprintf(sb, sizeof sb, "%08d", (int)b);
    return strcmp(sa, sb);
}
}  // namespace

// Union of per-shard tables without sorting anything (the rank-0 step of a multi-GPU run).  Shards are runs of
// consecutive contigs (plan_contig_shards), every shard's table arrives sorted by compare_junctions with names ranked
// inside the shard, and a coordinate-sorted BAM visits contigs in tid order: the global name of a junction is its local
// name plus the sizes of the shards before it, and the global order is a merge of the (already sorted) tables.  Every
// assumption is checked in O(n); anything else takes the general path below.
bool Engine::merge_sorted_shards() {
    if (!final_.empty() || import_sizes_.empty() || imported_.size() >= 100000000ull) return false;
    std::vector<uint32_t> cr = contig_ranks(contigs_);
    auto crank = [&](int32_t tid) { return (tid >= 0 && (size_t)tid < cr.size()) ? cr[(size_t)tid] : 0x40000000u + (uint32_t)tid; };
    auto bed_less = [&](const rtjx_junction& x, const rtjx_junction& y) {
        const uint32_t cx = crank(x.tid), cy = crank(y.tid);
        if (cx != cy) return cx < cy;
        if (x.thick_start != y.thick_start) return x.thick_start < y.thick_start;
        if (x.thick_end != y.thick_end) return x.thick_end < y.thick_end;
        return x.name_index < y.name_index;
    };
    struct Part { size_t lo, hi; int32_t tmin, tmax; };
    std::vector<Part> parts;
    size_t off = 0;
    for (size_t n : import_sizes_) { if (n) parts.push_back(Part{off, off + n, INT32_MAX, INT32_MIN}); off += n; }
    std::vector<const rtjx_junction*> by_name;
    for (Part& p : parts) {
        const size_t n = p.hi - p.lo;
        by_name.assign(n, nullptr);
        for (size_t i = p.lo; i < p.hi; ++i) {
            const rtjx_junction& j = imported_[i];
            if (j.name_index == 0 || j.name_index > n || by_name[j.name_index - 1]) return false;      // names: a permutation of 1..n
            by_name[j.name_index - 1] = &j;
            if (i > p.lo && bed_less(j, imported_[i - 1])) return false;                                 // sorted as delivered
            p.tmin = std::min(p.tmin, j.tid); p.tmax = std::max(p.tmax, j.tid);
        }
        for (size_t r = 1; r < n; ++r) {                                                               // name order == (tid, first seen)
            const rtjx_junction &a = *by_name[r - 1], &b = *by_name[r];
            if (a.tid > b.tid || (a.tid == b.tid && a.first_ord >= b.first_ord)) return false;
        }
    }
    std::sort(parts.begin(), parts.end(), [](const Part& a, const Part& b) { return a.tmin < b.tmin; });
    for (size_t k = 1; k < parts.size(); ++k) if (parts[k].tmin <= parts[k - 1].tmax) return false;     // disjoint runs of contigs
    uint32_t name_base = 0;
    final_.clear();
    final_.reserve(imported_.size());
    std::vector<rtjx_junction> tmp;
    for (const Part& p : parts) {
        const size_t mid = final_.size();
        for (size_t i = p.lo; i < p.hi; ++i) { rtjx_junction j = imported_[i]; j.name_index += name_base; final_.push_back(j); }
        name_base += (uint32_t)(p.hi - p.lo);
        if (mid) {
            tmp.resize(final_.size());
            std::merge(final_.begin(), final_.begin() + (long)mid, final_.begin() + (long)mid, final_.end(), tmp.begin(), bed_less);
            final_.swap(tmp);
        }
    }
    return true;
}

void Engine::host_rank_and_sort() {
    const bool trace = getenv("RTJX_TRACE") != nullptr;
    if (merge_sorted_shards()) { if (trace) fprintf(stderr, "[rtjx] shard merge: %zu tables merged without sorting\n", import_sizes_.size()); return; }
    if (trace && !imported_.empty()) fprintf(stderr, "[rtjx] shard merge: general path (re-rank + sort)\n");
    const bool sharded = !imported_.empty();
    final_.insert(final_.end(), imported_.begin(), imported_.end());
    std::vector<uint32_t> order(final_.size());
    std::iota(order.begin(), order.end(), 0u);
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        const rtjx_junction &x = final_[a], &y = final_[b];
        if (sharded && x.tid != y.tid) return x.tid < y.tid;
        return x.first_ord < y.first_ord;
    });
    for (size_t r = 0; r < order.size(); ++r) final_[order[r]].name_index = (uint32_t)(r + 1);
    std::vector<uint32_t> cr = contig_ranks(contigs_);
    auto crank = [&](int32_t tid) { return (tid >= 0 && (size_t)tid < cr.size()) ? cr[(size_t)tid] : 0x40000000u + (uint32_t)tid; };
    std::sort(final_.begin(), final_.end(), [&](const rtjx_junction& x, const rtjx_junction& y) {
        uint32_t cx = crank(x.tid), cy = crank(y.tid);
        if (cx != cy) return cx < cy;
        if (x.thick_start != y.thick_start) return x.thick_start < y.thick_start;
        if (x.thick_end != y.thick_end) return x.thick_end < y.thick_end;
        return name_cmp(x.name_index, y.name_index) < 0;
    });
}

std::vector<uint32_t> Engine::contig_rank_table() const { return contig_ranks(contigs_); }

int Engine::ensure_finalize_buffers(uint32_t n, size_t n_contigs) {
    if (n > fin_cap_) {
        CK(cudaDeviceSynchronize());
        cached_dev_free(d_out_); d_out_ = nullptr;
        uint32_t cap = std::max<uint32_t>(n + n / 2, 1u << 16);
        CK(cached_dev_malloc(&d_out_, (size_t)cap * sizeof(OutJunction)));
        fin_cap_ = cap;
        cached_dev_free(d_ws_); d_ws_ = nullptr;
        ws_cap_ = finalize_sort_workspace_bytes(cap);
        CK(cached_dev_malloc(&d_ws_, ws_cap_));
    }
    if (n > h_final_cap_) {
        cached_host_free(h_final_); h_final_ = nullptr;
        uint32_t cap = std::max<uint32_t>(n + n / 2, 1u << 16);
        CK(cached_host_alloc(&h_final_, (size_t)cap * sizeof(rtjx_junction)));
        h_final_cap_ = cap;
    }
    if (n_contigs > rank_cap_ || !d_rank_) {
        CK(cudaDeviceSynchronize());
        cached_dev_free(d_rank_); d_rank_ = nullptr;
        rank_cap_ = std::max<size_t>(n_contigs * 2, 64);
        CK(cached_dev_malloc(&d_rank_, rank_cap_ * 4));
        rank_dirty_ = true;
    }
    return RTJX_OK;
}

int Engine::finalize(cudaStream_t user_stream) {
    if (finalized_ && !dirty_) return RTJX_OK;
    NvtxRange nvtx("rtjx:finalize (compact + first-seen rank + sort + D2H)");
    final_.clear(); pinned_final_n_ = 0;
    uint32_t n = 0;
    if (dev_ready_ && d_table_) {
        cudaSetDevice(prm_.device);
        cudaStream_t st = user_stream ? user_stream : stream_;
        int rc = sync_counters(st);
        if (rc) return rc;
        n = h_counters_[CTR_NUNIQUE];
        TableRef tref = table_ref();
        bc_pairs_n_ = 0;
        if (bc_mode_ && n && (rc = fold_barcode_table(n, st, &tref, &n))) return rc;     // n: pairs -> junctions
        if (n) {
            if ((rc = ensure_finalize_buffers(n, contigs_.size()))) return rc;
            std::vector<uint32_t> cr;
            if (rank_dirty_) cr = contig_ranks(contigs_);
            const size_t ws_bytes = ws_cap_;
            if (rank_dirty_) {
                if (!cr.empty()) CK(cudaMemcpyAsync(d_rank_, cr.data(), cr.size() * 4, cudaMemcpyHostToDevice, st));
                CK(cudaStreamSynchronize(st));      // cr is a local
                rank_dirty_ = false;
            }
            cudaEvent_t ea = nullptr, eb = nullptr;
            if (prm_.profile) { ea = get_event(); eb = get_event(); cudaEventRecord(ea, st); }
            launch_table_compact(tref, n, d_out_, st);
            launch_finalize_sort(d_out_, n, d_rank_, (uint32_t)contigs_.size(), d_ws_, ws_bytes, st);
            if (prm_.profile) cudaEventRecord(eb, st);
            CK(cudaMemcpyAsync(h_final_, d_out_, (size_t)n * sizeof(OutJunction), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaGetLastError());
            pinned_final_n_ = n;                           // the table stays in the pinned buffer: count/get/print read it there
            if (prm_.profile) { float ms = 0; if (cudaEventElapsedTime(&ms, ea, eb) == cudaSuccess) stats_.finalize_ms += ms; ev_pool_.push_back(ea); ev_pool_.push_back(eb); }
            stats_.kernel_launches += 2;     // ours: compact + name assignment (CUB's merge-sort passes not counted)
            stats_.d2h_bytes += (size_t)n * sizeof(OutJunction);
        }
    }
    if (!imported_.empty() || n >= 100000000u) {
        if (pinned_final_n_) { final_.assign(h_final_, h_final_ + pinned_final_n_); pinned_final_n_ = 0; }
        host_rank_and_sort();
    }
    finalized_ = true; dirty_ = false;
    return RTJX_OK;
}

int64_t Engine::count() {
    int rc = finalize(nullptr);
    return rc ? rc : (int64_t)final_size();
}

int64_t Engine::get(rtjx_junction* out, size_t cap) {
    int rc = finalize(nullptr);
    if (rc) return rc;
    if (out && cap) memcpy(out, final_data(), std::min(cap, final_size()) * sizeof(rtjx_junction));
    return (int64_t)final_size();
}

int Engine::import(const rtjx_junction* j, size_t n) {
    if (n && !j) return fail(RTJX_E_ARG, "null junctions");
    imported_.insert(imported_.end(), j, j + n);
    import_sizes_.push_back(n);
    finalized_ = false;
    return RTJX_OK;
}

int Engine::clear() {
    const uint64_t known_unique = unique_upper_;      // upper bound of occupied slots
    final_.clear(); pinned_final_n_ = 0; imported_.clear(); import_sizes_.clear(); finalized_ = false; dirty_ = false; unique_upper_ = 0; next_ord_ = 0; run_ord_base_ = 0;
    bc_pairs_n_ = 0;
    bc_dict_.clear();                                  // the ids live in the table keys: both go together
    if (dev_ready_) {
        cudaSetDevice(prm_.device);
        if (d_table_) {
            // zero only the occupied slots when they are few, else the whole table
            if (known_unique < table_slots_ / 16) { launch_table_clear(table_ref(), d_counters_ + CTR_NUNIQUE, (uint32_t)known_unique + 1, stream_); stats_.kernel_launches++; }
            else CK(cudaMemsetAsync(d_table_, 0, (size_t)table_slots_ * sizeof(Slot), stream_));
        }
        CK(cudaMemsetAsync(d_counters_, 0, CTR_COUNT * sizeof(uint32_t), stream_));
        CK(cudaStreamSynchronize(stream_));
    }
    return RTJX_OK;
}

// ---- `-b` single-cell barcodes ------------------------------------------------------------------------------
// add_junction keeps an unordered_map<barcode, count> per junction (junctions_extractor.cc:203-215).  Here the device table
// is keyed (junction, barcode id): count = reads of that barcode, nfirst = first supporting N op.  The junction-level
// fields (count, thick ends, anchors, first/last ordinals) are associative reductions, so the table add_junction would
// have built is the fold of the pair table over the barcode bits of the key — done on the device into a second table that
// the normal compaction / naming / sort then reads.  The pairs travel to the host sorted by (junction, first ordinal).
int Engine::fold_barcode_table(uint32_t n_pairs, cudaStream_t st, TableRef* folded, uint32_t* n_junctions) {
    int rc;
    const uint32_t want = std::max<uint32_t>(next_pow2(2ull * n_pairs), 1u << 16);
    if (want > fold_slots_) {
        CK(cudaDeviceSynchronize());
        cached_dev_free(d_fold_table_); cached_dev_free(d_fold_list_); d_fold_table_ = nullptr; d_fold_list_ = nullptr;
        CK(cached_dev_malloc(&d_fold_table_, (size_t)want * sizeof(Slot)));
        CK(cached_dev_malloc(&d_fold_list_, (size_t)want * sizeof(uint32_t)));
        fold_slots_ = want;
    }
    if (!d_fold_counters_) CK(cached_dev_malloc(&d_fold_counters_, CTR_COUNT * sizeof(uint32_t)));
    if ((rc = ensure_region_buffers(n_pairs))) return rc;
    const TableRef dst{d_fold_table_, fold_slots_ - 1, d_fold_list_, fold_slots_};
    CK(cudaMemsetAsync(d_fold_table_, 0, (size_t)fold_slots_ * sizeof(Slot), st));
    CK(cudaMemsetAsync(d_fold_counters_, 0, CTR_COUNT * sizeof(uint32_t), st));
    launch_table_fold(table_ref(), n_pairs, dst, d_fold_counters_, st);
    launch_table_compact_regions(table_ref(), n_pairs, d_out_r_, st);
    launch_sort_barcode_pairs(d_out_r_, n_pairs, d_ws_r_, ws_r_cap_, st);
    uint32_t* hc = h_counters_ + CTR_COUNT;
    CK(cudaMemcpyAsync(hc, d_fold_counters_, CTR_COUNT * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h_final_r_, d_out_r_, (size_t)n_pairs * sizeof(OutJunctionR), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    stats_.kernel_launches += 2; stats_.d2h_bytes += (size_t)n_pairs * sizeof(OutJunctionR) + CTR_COUNT * sizeof(uint32_t);
    if (hc[CTR_CAND_OVERFLOW] || hc[CTR_NUNIQUE] > n_pairs) return fail(RTJX_E_STATE, "internal: barcode fold table overflowed");
    bc_pairs_n_ = n_pairs;
    *folded = dst;
    *n_junctions = hc[CTR_NUNIQUE];
    return RTJX_OK;
}

// print_barcodes (junctions_extractor.h:99-111) for every junction print_all_junctions prints (:267-273), in its order.
// The line lists an unordered_map in iteration order.  The reference's map went through one copy-assignment per
// supporting read (:208,:214); a copy keeps bucket count and node order, so its final order is that of ONE map into which
// the junction's distinct barcodes were inserted in first-seen order — replayed here with the same std::unordered_map.
int Engine::write_barcodes(int fd) {
    if (!bc_mode_) return fail(RTJX_E_STATE, "the handle was not created in -b mode (rtjx_params.barcode_out)");
    int rc = finalize(nullptr);
    if (rc) return rc;
    const OutJunctionR* pairs = h_final_r_;
    const uint32_t np = bc_pairs_n_;
    auto proxy_of = [](uint8_t c) -> uint32_t { return c == '+' ? 0u : (c == '-' ? 1u : 2u); };
    auto key_less = [&](const OutJunctionR& a, int32_t tid, uint32_t start, uint32_t end, uint32_t proxy) {
        if (a.j.tid != tid) return a.j.tid < tid;
        if (a.j.start != start) return a.j.start < start;
        if (a.j.end != end) return a.j.end < end;
        return proxy_of(a.j.strand) < proxy;
    };
    std::string out;
    out.reserve(1u << 20);
    auto flush = [&]() -> bool {
        size_t off = 0;
        while (off < out.size()) {
            ssize_t w = ::write(fd, out.data() + off, out.size() - off);
            if (w <= 0) return false;
            off += (size_t)w;
        }
        out.clear();
        return true;
    };
    const rtjx_junction* fin = final_data();
    for (size_t fi = 0, fn = final_size(); fi < fn; ++fi) {
        const rtjx_junction& j = fin[fi];
        if (!(j.left_ok && j.right_ok)) continue;
        const uint32_t proxy = proxy_of(j.strand);
        uint32_t lo = 0, hi = np;
        while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (key_less(pairs[mid], j.tid, j.start, j.end, proxy)) lo = mid + 1; else hi = mid; }
        std::unordered_map<std::string, int> m;
        uint64_t total = 0;
        for (uint32_t k = lo; k < np; ++k) {
            const OutJunctionR& r = pairs[k];
            if (r.j.tid != j.tid || r.j.start != j.start || r.j.end != j.end || proxy_of(r.j.strand) != proxy) break;
            if (r.region == 0 || r.region > bc_dict_.names.size()) return fail(RTJX_E_STATE, "internal: junction without a barcode");
            m.insert(std::pair<std::string, int>(bc_dict_.names[r.region - 1], (int)r.j.count));
            total += r.j.count;
        }
        if (total != j.read_count) return fail(RTJX_E_STATE, "internal: barcode counts do not add up to the junction's read count");
        out += std::to_string(m.size());
        out += '\t';
        for (std::unordered_map<std::string, int>::const_iterator it = m.begin(); it != m.end(); ++it) {
            if (it != m.begin()) out += ',';
            out += it->first; out += ':'; out += std::to_string(it->second);
        }
        out += '\n';
        if (out.size() > (1u << 20) && !flush()) return fail(RTJX_E_IO, "write failed");
    }
    if (!flush()) return fail(RTJX_E_IO, "write failed");
    return RTJX_OK;
}

int64_t Engine::intern_barcode(const char* s) {
    if (!bc_mode_) return fail(RTJX_E_STATE, "the handle was not created in -b mode (rtjx_params.barcode_out)");
    if (!s) return fail(RTJX_E_ARG, "null barcode");
    if (bc_dict_.names.size() >= (1u << 24) - 2u) return fail(RTJX_E_UNSUPPORTED, "more than 16.7 million distinct barcodes");
    return (int64_t)bc_dict_.intern(s, strlen(s));
}

int Engine::barcode_stats(uint64_t* n_barcodes, uint64_t* n_missing) {
    if (!bc_mode_) return fail(RTJX_E_STATE, "the handle was not created in -b mode (rtjx_params.barcode_out)");
    if (n_barcodes) *n_barcodes = bc_dict_.names.size();
    if (n_missing) *n_missing = bc_dict_.missing;
    return RTJX_OK;
}

const char* Engine::barcode_name(uint32_t id) {
    return id < bc_dict_.names.size() ? bc_dict_.names[id].c_str() : nullptr;
}

// ---- BED12 (Junction::print, junctions_extractor.h:90-98; anchor filter junctions_extractor.cc:267)
namespace {
inline char* put_u32(char* p, uint32_t v) {
    char tmp[10]; int n = 0;
    do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = tmp[--n];
    return p;
}
}  // namespace

int Engine::write_bed12(int fd) {
    int rc = finalize(nullptr);
    if (rc) return rc;
    NvtxRange nvtx("rtjx:write_bed12");
    // Junction::print (junctions_extractor.h:90-98) without per-line allocation or printf; large tables are formatted by a few
    // threads, each into its own buffer, and written in order (43 MB for the 100M-read BAM: 37 -> ~15 ms)
    std::vector<size_t> clen(contigs_.size());
    for (size_t i = 0; i < contigs_.size(); ++i) clen[i] = contigs_[i].size();
    size_t max_name = 8;
    for (size_t l : clen) max_name = std::max(max_name, l);
    const rtjx_junction* fin = final_data();
    const size_t fn = final_size();
    const size_t line_cap = max_name + 192;
    auto format_range = [&](size_t lo, size_t hi, std::vector<char>& out) {
        out.resize((hi - lo) * line_cap + 64);
        char* p = out.data();
        for (size_t fi = lo; fi < hi; ++fi) {
            const rtjx_junction& j = fin[fi];
            if (!(j.left_ok && j.right_ok)) continue;
            const bool known = j.tid >= 0 && (size_t)j.tid < contigs_.size();
            if (known) { memcpy(p, contigs_[(size_t)j.tid].data(), clen[(size_t)j.tid]); p += clen[(size_t)j.tid]; }
            else p += snprintf(p, 32, "tid%d", (int)j.tid);
            *p++ = '\t'; p = put_u32(p, j.thick_start);
            *p++ = '\t'; p = put_u32(p, j.thick_end);
            memcpy(p, "\tJUNC", 5); p += 5;
            {   // setfill('0') << setw(8) << int (junctions_extractor.cc:152-157)
                const int v = (int)j.name_index;
                if (v >= 0 && v < 100000000) { uint32_t x = (uint32_t)v; for (int d = 7; d >= 0; --d) { p[d] = (char)('0' + x % 10); x /= 10; } p += 8; }
                else { p += snprintf(p, 16, "%08d", v); }
            }
            *p++ = '\t'; p = put_u32(p, j.read_count);
            *p++ = '\t'; *p++ = (char)j.strand;
            *p++ = '\t'; p = put_u32(p, j.thick_start);
            *p++ = '\t'; p = put_u32(p, j.thick_end);
            memcpy(p, "\t255,0,0\t2\t", 11); p += 11;
            p = put_u32(p, j.start - j.thick_start); *p++ = ','; p = put_u32(p, j.thick_end - j.end);
            memcpy(p, "\t0,", 3); p += 3;
            p = put_u32(p, j.end - j.thick_start);
            *p++ = '\n';
        }
        out.resize((size_t)(p - out.data()));
    };
    auto write_all = [&](const std::vector<char>& v) -> bool {
        size_t off = 0;
        while (off < v.size()) {
            ssize_t w = ::write(fd, v.data() + off, v.size() - off);
            if (w <= 0) return false;
            off += (size_t)w;
        }
        return true;
    };
    const size_t CHUNK = 32768;                                  // junctions per formatting task
    const int hw = prm_.n_threads > 0 ? prm_.n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    const int T = fn >= 4 * CHUNK ? std::min(8, hw) : 1;
    if (T <= 1) {
        std::vector<char> out;
        for (size_t lo = 0; lo < fn; lo += CHUNK) {
            format_range(lo, std::min(fn, lo + CHUNK), out);
            if (!write_all(out)) return fail(RTJX_E_IO, "write failed");
        }
        return RTJX_OK;
    }
    const size_t n_chunks = (fn + CHUNK - 1) / CHUNK;
    std::vector<std::vector<char>> outs(n_chunks);
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < T; ++t)
        pool.emplace_back([&] {
            for (size_t c; (c = next.fetch_add(1)) < n_chunks;) format_range(c * CHUNK, std::min(fn, (c + 1) * CHUNK), outs[c]);
        });
    for (auto& th : pool) th.join();
    for (const auto& o : outs) if (!write_all(o)) return fail(RTJX_E_IO, "write failed");
    return RTJX_OK;
}

const char* Engine::contig(int32_t tid) {
    if (tid >= 0 && (size_t)tid < contigs_.size()) return contigs_[(size_t)tid].c_str();
    unknown_contig_ = "tid" + std::to_string(tid);
    return unknown_contig_.c_str();
}

int32_t Engine::intern_contig(const char* name) {
    if (!name) return -1;
    for (size_t i = 0; i < contigs_.size(); ++i) if (contigs_[i] == name) return (int32_t)i;
    contigs_.push_back(name);
    finalized_ = false; rank_dirty_ = true;
    return (int32_t)contigs_.size() - 1;
}

void Engine::get_stats(rtjx_stats* out) {
    if (dev_ready_) {
        cudaSetDevice(prm_.device);
        resolve_profile_events();
    }
    stats_.table_slots = table_slots_;
    *out = stats_;
}

void Engine::reset_stats() {
    if (dev_ready_) { cudaSetDevice(prm_.device); resolve_profile_events(); }
    memset(&stats_, 0, sizeof stats_);
}

}  // namespace rtjx
