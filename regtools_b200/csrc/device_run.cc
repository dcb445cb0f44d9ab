// regtools_b200/csrc/device_run.cc — whole-file runs with BGZF inflate and BAM record split on the device.
//
// Host work per chunk of the file: walk BGZF block headers (18-byte header, BSIZE, ISIZE trailer —
// bgzf.c:348-355,525-546), copy the compressed bytes from the page cache into pinned memory with a
// few threads, enqueue.  Everything else (inflate, record-start discovery, record chain walk, field extraction,
// CIGAR scan, junction merge) runs on the GPU.
//
// Round-2 shape of the pipeline (round 1: one stream, one group at a time, scan per group):
//   * a GROUP of consecutive BGZF blocks is the unit of device work; up to eight group slots are in flight: the inflates of the
//     groups behind g (each on its slot's stream; the lane-per-stream decoder is latency-bound, so concurrent launches add up)
//     overlap the record walk / extraction of group g (the chain stream), and all overlap the H2D copies (copy stream);
//   * record starts are found ON the device, one per BGZF block (block_seeds_kernel), instead of every 16 kb of reference
//     from the index — segments of <= 64 KiB instead of hundreds of thousands of records on deep loci; the chain walk verifies
//     every guessed start, a miss makes the run start over with the index seeds of round 1 (then the linear index alone,
//     then the host feeder), so exactness never depends on a guess;
//   * extracted alignments are APPENDED to one SoA in HBM across groups; cigar_scan + junction_merge run once per tens of
//     millions of alignments (or at the end), so the scan kernel works on batches large enough to stream at its full rate;
//   * the compressed file can be staged in HBM beforehand (rtjx_stage_bam): the run then reads it from there and does no
//     host copy at all — that is the "input resident in HBM" configuration bench.py reports as `value`.
// Any anomaly the device reports (malformed record, a walk that misses its next seed, capacity) makes
// Engine::run fall back, so results never depend on this path accepting bad input.
#include "engine.h"
#include "buffer_cache.h"
#include "stage_pipe.h"

#include <algorithm>
#include <chrono>
#include <atomic>
#include <climits>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <unistd.h>

namespace rtjx {

namespace {
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// page cache -> pinned memory with pread (no page-table faults, unlike memcpy from the mmap), several threads
size_t parallel_pread(int fd, uint8_t* dst, size_t n, uint64_t file_off, int threads) {
    auto one = [fd](uint8_t* d, size_t len, uint64_t off) -> size_t {
        size_t got = 0;
        while (got < len) {
            ssize_t r = pread(fd, d + got, len - got, (off_t)(off + got));
            if (r <= 0) break;
            got += (size_t)r;
        }
        return got;
    };
    if (n < (4u << 20) || threads <= 1) return one(dst, n, file_off);
    std::vector<std::thread> pool;
    std::vector<size_t> got((size_t)threads, 0);
    const size_t per = ((n + threads - 1) / threads + 4095) & ~(size_t)4095;
    int used = 0;
    for (int t = 0; t < threads; ++t) {
        const size_t o = (size_t)t * per;
        if (o >= n) break;
        const size_t len = std::min(per, n - o);
        ++used;
        pool.emplace_back([&, t, o, len] { got[(size_t)t] = one(dst + o, len, file_off + o); });
    }
    for (auto& th : pool) th.join();
    size_t total = 0;
    for (int t = 0; t < used; ++t) { total += got[(size_t)t]; if (got[(size_t)t] < std::min(per, n - (size_t)t * per)) break; }
    return total;
}

}  // namespace

#define CKD(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return fail(RTJX_E_CUDA, std::string(#call " failed: ") + cudaGetErrorString(e__));      \
    } while (0)

struct Engine::DeviceFeed {
    static constexpr uint32_t HEAD = 4u << 20;           // carry headroom in front of the inflated data
    static constexpr int MAXSLOT = 16;
    int NSLOT = 8;                                       // groups in flight (their inflates run concurrently, each on its slot's stream:
                                                         // a 384 MB group occupies ~2 of the 16 warps per SM the lane decoder can hold);
                                                         // RTJX_FEED_SLOTS overrides.  Measured on the 100M-read file, resident pass:
                                                         // 256 MB x 10 slots 386 ms, 384 x 8 330, 512 x 6 332, 512 x 10 320 (twice the
                                                         // buffers), 1024 x 5 379, 128 x 16 675: a launch ends with its slowest warp, and
                                                         // the fewer, larger launches there are, the less of the GPU waits on such tails
    static constexpr int NSTAGE = StagePipe::NBUF;       // pinned staging windows (read ahead by StagePipe)
    uint8_t* h_comp[NSTAGE] = {};
    cudaEvent_t comp_free[NSTAGE] = {};
    struct GroupSlot {
        uint8_t* d_comp = nullptr; size_t comp_cap = 0;  // compressed bytes of the group (file mode)
        uint8_t* d_infl = nullptr; size_t infl_cap = 0;  // HEAD + data + pad
        BgzfBlockDesc* d_desc = nullptr; uint32_t* d_status = nullptr; size_t desc_cap = 0;
        void* d_scratch = nullptr; size_t scratch_cap = 0;
        int64_t* d_seeds = nullptr; uint32_t* d_segbase = nullptr; uint32_t* d_segcnt = nullptr; uint32_t* d_segscan = nullptr; size_t seed_cap = 0;
        int32_t* d_recoff = nullptr; int32_t* d_dense = nullptr; uint32_t* d_ncig = nullptr; uint32_t* d_ncigscan = nullptr; size_t rec_cap = 0;
        void* d_ws = nullptr; size_t ws_cap = 0;
        FeedState* h_state = nullptr;                    // pinned copy of the device state after this slot's group
        cudaStream_t inf_stream = nullptr;               // inflate + record-start discovery of this slot's group; the engine's stream_ runs the chain
        cudaEvent_t copied = nullptr, inflated = nullptr, done = nullptr;
        bool busy = false;                               // a group was launched into this slot and not yet checked
        uint64_t seq = 0, cap_rec = 0, cap_ops = 0;      // launch number; upper bounds of what the group appends to the accumulator
        uint32_t dbg_nb = 0, dbg_nseg = 0; uint64_t dbg_out_total = 0; bool dbg_first = false;   // RTJX_FEED_DEBUG
    } slot[MAXSLOT];
    FeedState* d_state = nullptr;
    uint8_t* d_carry = nullptr;                          // HEAD bytes: the unfinished record at the end of a group, right-aligned
    // SoA accumulator: the alignments extracted since the last cigar_scan
    int32_t* a_tid = nullptr; int32_t* a_pos = nullptr; uint32_t* a_meta = nullptr; uint32_t* a_off = nullptr; size_t acc_rec_cap = 0;
    uint32_t* a_cigar = nullptr; size_t acc_ops_cap = 0;
    // compressed file staged in HBM (rtjx_stage_bam); the BGZF header walk of every chunk of it is remembered (the file does
    // not change under a staged handle), so later runs do no host pass over the file at all
    uint8_t* d_file = nullptr; size_t file_bytes = 0;
    struct ChunkScan { std::vector<BgzfBlockInfo> blocks; uint64_t c_end; uint64_t end_coff; size_t got; bool stop, partial, untrusted; };
    std::unordered_map<uint64_t, ChunkScan> scans;

    ~DeviceFeed() {
        for (int i = 0; i < NSTAGE; ++i) { cached_host_free(h_comp[i]); if (comp_free[i]) cudaEventDestroy(comp_free[i]); }
        for (GroupSlot& s : slot) {
            cached_dev_free(s.d_comp); cached_dev_free(s.d_infl); cached_dev_free(s.d_desc); cached_dev_free(s.d_status); cached_dev_free(s.d_scratch);
            cached_dev_free(s.d_seeds); cached_dev_free(s.d_segbase); cached_dev_free(s.d_segcnt); cached_dev_free(s.d_segscan);
            cached_dev_free(s.d_recoff); cached_dev_free(s.d_dense); cached_dev_free(s.d_ncig); cached_dev_free(s.d_ncigscan); cached_dev_free(s.d_ws);
            cached_host_free(s.h_state);
            if (s.copied) cudaEventDestroy(s.copied);
            if (s.inflated) cudaEventDestroy(s.inflated);
            if (s.done) cudaEventDestroy(s.done);
            if (s.inf_stream) cudaStreamDestroy(s.inf_stream);
        }
        cached_dev_free(d_state); cached_dev_free(d_carry);
        cached_dev_free(a_tid); cached_dev_free(a_pos); cached_dev_free(a_meta); cached_dev_free(a_off); cached_dev_free(a_cigar);
        cached_dev_free(d_file);
    }
};

void Engine::DeviceFeedDeleter::operator()(DeviceFeed* p) const { delete p; }

// One group = the unit of device work: a run of consecutive BGZF blocks (several staging chunks) that is
// inflated, split and appended to the alignment accumulator by one set of kernel launches.
struct FeedGroup {
    std::vector<BgzfBlockDesc> desc;     // in_off relative to the group's first byte
    std::vector<uint64_t> coffs;         // compressed offset of every block (for seed lookup)
    std::vector<uint32_t> isize;
    uint64_t comp_bytes = 0, out_total = 0, first_coff = 0;
    uint64_t v_lo = 0;                   // first virtual offset of the group (record start for the first group of a range)
    bool first_of_range = false;
    void clear() { desc.clear(); coffs.clear(); isize.clear(); comp_bytes = out_total = 0; first_coff = 0; }
};

// rtjx_stage_bam: the compressed file goes to HBM once; later runs of this handle read it from there.
int Engine::stage_file() {
    int rc = ensure_device();
    if (rc) return rc;
    BamFile bam; std::string err;
    if (bam_path_.empty() || !bam.open(bam_path_, &err)) return fail(RTJX_E_OPEN_BAM, "Unable to open BAM/SAM file.\n\n");
    if (!dfeed_) dfeed_.reset(new DeviceFeed());
    DeviceFeed& F = *dfeed_;
    CKD(cudaDeviceSynchronize());
    cached_dev_free(F.d_file); F.d_file = nullptr; F.file_bytes = 0; F.scans.clear();
    CKD(cached_dev_malloc(&F.d_file, bam.size() + 256));
    const size_t CH = 64u << 20;
    uint8_t* h = nullptr;
    CKD(cached_host_alloc(&h, CH));
    const int threads = prm_.n_threads > 0 ? prm_.n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    for (size_t o = 0; o < bam.size(); o += CH) {
        const size_t n = std::min(CH, bam.size() - o);
        if (parallel_pread(bam.fd(), h, n, o, threads) != n) { cached_host_free(h); return fail(RTJX_E_IO, "short read while staging the BAM"); }
        CKD(cudaMemcpy(F.d_file + o, h, n, cudaMemcpyHostToDevice));
    }
    CKD(cudaMemset(F.d_file + bam.size(), 0, 256));
    cached_host_free(h);
    F.file_bytes = bam.size();
    stats_.h2d_bytes += bam.size();
    return RTJX_OK;
}

// Returns RTJX_OK, a negative status, or +1 = "device path declined, use the next seed mode / the host feeder".
int Engine::run_device(const BamFile& bam, const BaiIndex& idx, const IterSpec& spec) {
    NvtxRange nvtx_run("rtjx:device_feed");
    const double t_begin = now_s();
    int rc = ensure_device();
    if (rc) return rc;
    if (!dfeed_) dfeed_.reset(new DeviceFeed());
    DeviceFeed& F = *dfeed_;
    const bool resident = F.d_file != nullptr && F.file_bytes == bam.size();
    const int n_threads = prm_.n_threads > 0 ? prm_.n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    const int copy_threads = std::min(n_threads, 16);
    const uint64_t STAGE = 16ull << 20;                  // compressed bytes per pinned staging window (plus WIN_SLACK: the block that straddles its end)
    const uint64_t WIN_SLACK = 128u << 10;
    static const uint64_t GROUP = [] { const char* v = getenv("RTJX_GROUP_MB"); return (uint64_t)(v ? atoi(v) : 384) << 20; }();
    // the first group of a range is smaller: the GPU starts after a few ms of staging instead of a full group's worth
    static const uint64_t FIRST_GROUP = [] { const char* v = getenv("RTJX_FIRST_GROUP_MB"); return (uint64_t)(v ? atoi(v) : 128) << 20; }();
    // alignments (upper bound, 64 bytes of stream each) accumulated before cigar_scan runs
    static const uint64_t ACC_REC = [] { const char* v = getenv("RTJX_SCAN_BATCH_M"); return (uint64_t)(v ? atoi(v) : 192) << 20; }();
    const uint64_t OUT_CAP = 1280ull << 20;              // inflated bytes at which a group is closed whatever its compressed size
    const int seed_mode = feed_seed_mode_;               // 0 device-found starts, 1 index (linear + bin chunks), 2 linear index only

    // ---- ranges to stream (same as the host feeder)
    std::vector<Chunk64> ranges;
    if (spec.kind == IterSpec::WholeFile) {
        uint64_t off0;
        if (!idx.whole_file_start(&off0)) return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
        if (off0 == 0) off0 = bam.header().first_record_voffset;
        ranges.push_back(Chunk64{off0, UINT64_MAX});
    } else if (spec.kind == IterSpec::Contigs) {
        ranges = coalesced_contig_ranges(idx, spec.contigs);   // a shard of consecutive contigs streams as one range
    } else if (spec.kind == IterSpec::Region) {
        // (Engine::run_impl: only with the one-region filter installed) first to last chunk of hts_itr_query, as one byte span
        const std::vector<Chunk64> off = idx.query(spec.tid, spec.beg, spec.end);
        if (off.empty()) return RTJX_OK;
        ranges.push_back(Chunk64{off.front().beg, off.back().end});
    } else {
        return 1;
    }
    // ---- record-start virtual offsets known to the index (seed modes 1 and 2)
    std::vector<uint64_t> seeds_all;
    if (seed_mode != 0) {
        // (the 16 kb linear index: one record start per window that holds reads; already in file order for a sorted BAM)
        for (const BaiIndex::Ref& r : idx.refs)
            for (uint64_t v : r.ioffset) if (v && (seeds_all.empty() || v != seeds_all.back())) seeds_all.push_back(v);
        // (the binning index: every chunk begins at a record and ends right after one (hts_idx_push, hts.c:1288-1350).  The
        // pseudo-bin's second "chunk" holds counts, not offsets.)
        if (seed_mode == 1)
            for (const BaiIndex::Ref& r : idx.refs)
                for (const BaiIndex::Bin& b : r.bins) {
                    if (b.bin == idx.meta_bin()) continue;
                    for (const Chunk64& c : b.chunks) { if (c.beg) seeds_all.push_back(c.beg); if (c.end) seeds_all.push_back(c.end); }
                }
        if (!std::is_sorted(seeds_all.begin(), seeds_all.end())) std::sort(seeds_all.begin(), seeds_all.end());
        seeds_all.erase(std::unique(seeds_all.begin(), seeds_all.end()), seeds_all.end());
    }

    // ---- fixed-size device state, pinned staging, streams
    if (!F.d_state) {
        if (const char* v = getenv("RTJX_FEED_SLOTS")) F.NSLOT = std::min(std::max(atoi(v), 2), (int)DeviceFeed::MAXSLOT);
        // the inflate streams run at the lowest priority, the engine's stream (record split, scan, merge) at the highest: the chain of a
        // finished group must not queue behind the long-running decoder CTAs of the groups after it
        int prio_lo = 0, prio_hi = 0;
        CKD(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CKD(cached_dev_malloc(&F.d_state, sizeof(FeedState)));
        CKD(cached_dev_malloc(&F.d_carry, (size_t)DeviceFeed::HEAD + 256));
        for (DeviceFeed::GroupSlot& s : F.slot) {
            CKD(cudaStreamCreateWithPriority(&s.inf_stream, cudaStreamNonBlocking, prio_lo));
            CKD(cached_host_alloc(&s.h_state, sizeof(FeedState)));
            CKD(cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming));
            CKD(cudaEventCreateWithFlags(&s.inflated, cudaEventDisableTiming));
            CKD(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
        }
    }
    if (!resident && !F.h_comp[0])
        for (int i = 0; i < DeviceFeed::NSTAGE; ++i) {
            CKD(cudaEventCreateWithFlags(&F.comp_free[i], cudaEventDisableTiming));
            CKD(cached_host_alloc(&F.h_comp[i], STAGE + WIN_SLACK + 512));
        }
    auto drain = [&]() { cudaStreamSynchronize(stream_); cudaStreamSynchronize(copy_stream_); for (DeviceFeed::GroupSlot& q : F.slot) cudaStreamSynchronize(q.inf_stream); };
    auto grow_dev = [&](void** p, size_t* cap, size_t want, size_t elem) -> cudaError_t {
        if (want <= *cap) return cudaSuccess;
        if (*p) drain();                                 // only a live buffer needs the streams drained
        cached_dev_free(*p); *p = nullptr;
        size_t c = want + want / 8;
        cudaError_t e = cached_dev_malloc(p, c * elem + 256);
        if (e == cudaSuccess) *cap = c;
        return e;
    };
    for (DeviceFeed::GroupSlot& s : F.slot) s.busy = false;
    std::unique_ptr<StagePipe> pipe;
    if (!resident) pipe.reset(new StagePipe(bam.fd(), bam.size(), F.h_comp, [&F](int b) { cudaEventSynchronize(F.comp_free[b]); }, STAGE + WIN_SLACK, STAGE, copy_threads));
    launch_feed_acc_reset(F.d_state, stream_);           // a declined earlier run may have left alignments in the accumulator

    static const bool trace = getenv("RTJX_TRACE") != nullptr;
    double t_stage = 0, t_launch = 0, t_scanhdr = 0, t_seeds = now_s() - t_begin, t_slot = 0, t_alloc = 0;
    const int xs_mode = prm_.strandness == 0;
    const int32_t n_ref = (int32_t)bam.header().names.size();
    uint64_t ordinal = run_ord_base_;                    // ordinal of the next alignment handed to cigar_scan
    // What the accumulator holds, as far as the host knows: the exact counts after the newest group it has seen finish, plus upper
    // bounds (64 B of stream per alignment, 32 B per CIGAR op) for the groups still in flight
    uint64_t known_rec = 0, known_ops = 0, known_seq = 0, infl_rec = 0, infl_ops = 0, launch_seq = 0;
    bool declined = false, reached_limit = false, acc_dirty = false;
    int n_groups = 0;

    // check(): waits for the group in `s` and looks at the state it left
    auto check = [&](DeviceFeed::GroupSlot& s) -> int {
        if (!s.busy) return 0;
        s.busy = false;
        const double tw = now_s();
        CKD(cudaEventSynchronize(s.done));
        stats_.host_wait_s += now_s() - tw;
        stats_.d2h_bytes += sizeof(FeedState);
        const FeedState& st = *s.h_state;
        infl_rec -= s.cap_rec; infl_ops -= s.cap_ops;
        if (s.seq > known_seq) { known_seq = s.seq; known_rec = st.acc_rec; known_ops = st.acc_ops; }
        if (st.flags || st.bad_offset != LLONG_MAX) {
            declined = true; feed_decline_flags_ = st.flags;
            if (getenv("RTJX_FEED_DEBUG") && s.dbg_nb) {
                // developer aid: which seed did the chain miss?  Walks the group's inflated stream on the host.
                drain();
                std::vector<int64_t> seeds(s.dbg_nseg);
                std::vector<uint8_t> data((size_t)s.dbg_out_total);
                std::vector<BgzfBlockDesc> desc(s.dbg_nb);
                cudaMemcpy(seeds.data(), s.d_seeds, seeds.size() * 8, cudaMemcpyDeviceToHost);
                cudaMemcpy(data.data(), s.d_infl + DeviceFeed::HEAD, data.size(), cudaMemcpyDeviceToHost);
                cudaMemcpy(desc.data(), s.d_desc, desc.size() * sizeof(BgzfBlockDesc), cudaMemcpyDeviceToHost);
                fprintf(stderr, "[rtjx dbg] group declined: flags %u bad_offset %lld n_seg %u out_total %llu first_of_range %d carry_len %u\n", st.flags,
                        st.bad_offset, s.dbg_nseg, (unsigned long long)s.dbg_out_total, (int)s.dbg_first, st.carry_len);
                if (seeds.size() > 2) {
                    int64_t p = s.dbg_first ? seeds[0] : seeds[1]; size_t si = s.dbg_first ? 1 : 2; int shown = 0;
                    while (p + 4 <= (int64_t)data.size() && shown < 8) {
                        int32_t bl; memcpy(&bl, &data[(size_t)p], 4);
                        while (si < seeds.size() && seeds[si] < p) {
                            fprintf(stderr, "[rtjx dbg]   seed %zu = %lld is NOT a record start (chain passed it at %lld; block out_off %u len %u)\n", si,
                                    (long long)seeds[si], (long long)p, desc[si].out_off, desc[si].out_len);
                            ++si; ++shown;
                        }
                        if (si < seeds.size() && seeds[si] == p) ++si;
                        if (bl < 32) { fprintf(stderr, "[rtjx dbg]   bad block_size %d at %lld\n", bl, (long long)p); break; }
                        p += 4 + (int64_t)bl;
                    }
                    fprintf(stderr, "[rtjx dbg]   host chain ended at %lld of %zu; seeds matched up to %zu of %zu\n", (long long)p, data.size(), si, seeds.size());
                    for (size_t k = 1; k < seeds.size() && k < 4; ++k) fprintf(stderr, "[rtjx dbg]   seed[%zu] = %lld (block out_off %u)\n", k, (long long)seeds[k], desc[k].out_off);
                    for (size_t k = seeds.size() > 3 ? seeds.size() - 3 : 1; k < seeds.size(); ++k) fprintf(stderr, "[rtjx dbg]   seed[%zu] = %lld (block out_off %u len %u)\n", k, (long long)seeds[k], desc[k].out_off, desc[k].out_len);
                }
            }
        }
        if (st.reached_limit) reached_limit = true;
        return 0;
    };
    // flush(): cigar_scan + junction_merge over what the accumulator holds
    auto flush = [&]() -> int {
        NvtxRange nvtx("rtjx:flush accumulator (wait + scan + merge)");
        for (DeviceFeed::GroupSlot& s : F.slot) { int r = check(s); if (r) return r; }
        if (declined || !acc_dirty) return 0;
        CKD(cudaStreamSynchronize(stream_));
        FeedState st;
        CKD(cudaMemcpy(&st, F.d_state, sizeof st, cudaMemcpyDeviceToHost));
        if (st.acc_rec) {
            BatchView v;
            v.n_reads = st.acc_rec; v.n_ops = st.acc_ops; v.first_ordinal = ordinal;
            v.tid = F.a_tid; v.pos = F.a_pos; v.meta = F.a_meta; v.cig_off = F.a_off; v.cigar = F.a_cigar;
            if (st.acc_jops) { int r = process_device_batch(v, st.acc_jops, stream_); if (r) return r; }
            else { stats_.reads += st.acc_rec; stats_.cigar_ops += st.acc_ops; stats_.batches++; }
            ordinal += st.acc_rec;
        }
        launch_feed_acc_reset(F.d_state, stream_);
        stats_.kernel_launches++;
        known_rec = known_ops = infl_rec = infl_ops = 0; known_seq = launch_seq; acc_dirty = false;
        return 0;
    };

    for (size_t ri = 0; ri < ranges.size() && !declined; ++ri) {
        const Chunk64 rg = ranges[ri];
        uint64_t coff = rg.beg >> 16;
        const bool bounded = rg.end != UINT64_MAX;
        // blocks needed: up to the one holding the end offset (exclusive if the end sits on a block boundary)
        const uint64_t end_coff = std::min<uint64_t>(bounded ? ((rg.end & 0xffff) ? (rg.end >> 16) + 1 : (rg.end >> 16)) : bam.size(), bam.size());
        bool first_group = true, stream_ends = false;
        reached_limit = false;
        int buf = 0;
        uint64_t win = coff;                                 // first byte of the staging window being consumed (file mode)
        FeedGroup g;
        g.first_of_range = true; g.v_lo = rg.beg;

        // launch(G, S): inflate + record starts on the inflate stream; walk + gather + extract + carry on the chain stream
        auto launch = [&](FeedGroup& G, DeviceFeed::GroupSlot& S) -> int {
            NvtxRange nvtx("rtjx:group (inflate | record starts | walk + gather + extract) enqueue");
            const double tl0 = now_s();
            struct TL { double* acc; double t0; ~TL() { *acc += now_s() - t0; } } tl{&t_launch, tl0};
            const uint32_t nb = (uint32_t)G.desc.size();
            // segments: one per BGZF block (seed mode 0), or one per record start the index knows inside this group
            std::vector<int64_t> sd;
            sd.push_back(G.first_of_range ? (int64_t)(G.v_lo & 0xffff) : 0);
            if (seed_mode == 0) {
                for (uint32_t i = 1; i < nb; ++i) sd.push_back((int64_t)G.desc[i].out_off);     // replaced on the device
            } else {
                const uint64_t v_lo = G.first_of_range ? G.v_lo : (G.coffs.front() << 16);
                const uint64_t v_hi = (G.coffs.back() << 16) + 0x10000;
                size_t bi = 0;
                for (auto it = std::upper_bound(seeds_all.begin(), seeds_all.end(), v_lo); it != seeds_all.end() && *it < v_hi; ++it) {
                    if (bounded && *it >= rg.end) break;
                    const uint64_t sc = *it >> 16, su = *it & 0xffff;
                    while (bi < nb && G.coffs[bi] < sc) ++bi;
                    if (bi == nb) break;
                    if (G.coffs[bi] != sc || su > G.isize[bi]) continue;        // not a block of this file
                    const int64_t o = (int64_t)G.desc[bi].out_off + (int64_t)su;
                    if (o > sd.back()) sd.push_back(o);
                }
            }
            const uint32_t n_seg = (uint32_t)sd.size();
            std::vector<uint32_t> segbase(n_seg + 1);
            uint64_t cap_total = 0;
            for (uint32_t i = 0; i < n_seg; ++i) {
                segbase[i] = (uint32_t)cap_total;
                const int64_t hi = i + 1 < n_seg ? sd[i + 1] : (int64_t)G.out_total;
                int64_t span = hi - sd[i];
                if (i == 0 && !G.first_of_range) span += DeviceFeed::HEAD;          // segment 0 starts inside the carry
                cap_total += (uint64_t)(span / 64 + 4);     // records are >= 37 bytes; denser than 64 B/record -> capacity flag -> fall back
            }
            segbase[n_seg] = (uint32_t)cap_total;
            // (record offsets inside a group are 31-bit: a group of a very compressible file that still got too large is left to the
            // host feeder — groups are closed at OUT_CAP inflated bytes below, so this takes a chunk that inflates 50-fold)
            if (cap_total > 0x7ffffff0ull || G.out_total > 0x7ff00000ull) { declined = true; feed_decline_flags_ = 0x80000000u; return 0; }
            // range end inside this group?
            int64_t limit = LLONG_MAX;
            if (bounded) {
                const uint64_t ec = rg.end >> 16, eu = rg.end & 0xffff;
                size_t k = std::lower_bound(G.coffs.begin(), G.coffs.end(), ec) - G.coffs.begin();
                if (k < nb && G.coffs[k] == ec) limit = (int64_t)G.desc[k].out_off + (int64_t)eu;
                // (the range ends exactly where this group's last block ends: only the group whose compressed bytes reach `ec` — any
                // earlier group also has all its blocks in front of `ec`, and a walk that happened to end on its last byte would
                // report the range finished: shard 1 of 4 of the 100M-read file stopped after 78k of its 119k blocks that way)
                else if (k == nb && eu == 0 && G.first_coff + G.comp_bytes == ec) limit = (int64_t)G.out_total;
            }
            // ---- device buffers of the slot and the accumulator.  A range's first group is a small one: its slot's buffers are
            // sized for the full-size group that slot will see next, so nothing is reallocated (and no stream drained) mid-run
            const double ta0 = now_s();
            // (a full-size group: GROUP compressed bytes, or OUT_CAP inflated ones — whichever closes it first — plus the chunk that
            // crosses the line)
            const double ratio = (double)std::max<uint64_t>(G.out_total, 1) / (double)std::max<uint64_t>(G.comp_bytes, 1);
            const double full_out = std::min((double)(GROUP + STAGE) * ratio, (double)OUT_CAP + (double)STAGE * ratio);
            const double grow = std::min(8.0, std::max(1.0, 1.05 * full_out / (double)std::max<uint64_t>(G.out_total, 1)));
            auto sized = [&](size_t need) -> size_t { return (size_t)((double)need * grow) + 64; };
            const uint64_t cig_upper = (DeviceFeed::HEAD + G.out_total) / 32 + 4096;   // > 12.5 % of the bytes being CIGAR -> capacity flag
            auto acc_full = [&] { return known_rec + infl_rec + cap_total + 8 > F.acc_rec_cap || known_ops + infl_ops + cig_upper + 8 > F.acc_ops_cap; };
            if (acc_full()) {
                // groups that have finished tighten the bound: look at them (oldest first) before giving up on the room
                for (int k = 0; k < F.NSLOT && acc_full(); ++k) {
                    int r = check(F.slot[(n_groups + k) % F.NSLOT]);
                    if (r) return r;
                    if (declined) return 0;
                }
            }
            if (acc_full()) {
                // the accumulator is scanned (and emptied) before it would overflow; it only ever grows while empty
                if (acc_dirty) { int r = flush(); if (r) return r; if (declined) return 0; }
                const uint64_t want_rec = std::max<uint64_t>(std::min<uint64_t>(ACC_REC, (uint64_t)(bam.size() / 16)), cap_total + 8);
                if (want_rec > F.acc_rec_cap) {
                    drain();
                    cached_dev_free(F.a_tid); cached_dev_free(F.a_pos); cached_dev_free(F.a_meta); cached_dev_free(F.a_off);
                    F.a_tid = F.a_pos = nullptr; F.a_meta = F.a_off = nullptr;
                    const size_t cap = want_rec + want_rec / 8;
                    CKD(cached_dev_malloc(&F.a_tid, cap * 4 + 256)); CKD(cached_dev_malloc(&F.a_pos, cap * 4 + 256));
                    CKD(cached_dev_malloc(&F.a_meta, cap * 4 + 256)); CKD(cached_dev_malloc(&F.a_off, (cap + 8) * 4 + 256));
                    F.acc_rec_cap = cap;
                }
                const uint64_t want_ops = std::max<uint64_t>(F.acc_rec_cap * 2, cig_upper + 8);
                if (want_ops > F.acc_ops_cap) {
                    drain();
                    cached_dev_free(F.a_cigar); F.a_cigar = nullptr;
                    const size_t cap = want_ops + want_ops / 8;
                    CKD(cached_dev_malloc(&F.a_cigar, (cap + 8) * 4 + 256));
                    F.acc_ops_cap = cap;
                }
            }
            if (DeviceFeed::HEAD + G.out_total + 64 > S.infl_cap) {
                if (S.d_infl) drain();
                cached_dev_free(S.d_infl); S.d_infl = nullptr;
                const size_t want = DeviceFeed::HEAD + sized(G.out_total);
                const size_t cap = want + want / 8;
                CKD(cached_dev_malloc(&S.d_infl, cap + 256));
                S.infl_cap = cap;
            }
            if ((size_t)nb + 2 > S.desc_cap) {
                if (S.d_desc) drain();
                cached_dev_free(S.d_desc); cached_dev_free(S.d_status); S.d_desc = nullptr; S.d_status = nullptr;
                S.desc_cap = sized((size_t)nb) + 1024;
                CKD(cached_dev_malloc(&S.d_desc, S.desc_cap * sizeof(BgzfBlockDesc))); CKD(cached_dev_malloc(&S.d_status, S.desc_cap * 4));
            }
            if (bgzf_inflate_scratch_bytes(nb) > S.scratch_cap) CKD(grow_dev(&S.d_scratch, &S.scratch_cap, bgzf_inflate_scratch_bytes((uint32_t)sized(nb)), 1));
            if ((size_t)n_seg + 2 > S.seed_cap) {
                if (S.d_seeds) drain();
                cached_dev_free(S.d_seeds); cached_dev_free(S.d_segbase); cached_dev_free(S.d_segcnt); cached_dev_free(S.d_segscan);
                S.d_seeds = nullptr; S.d_segbase = S.d_segcnt = S.d_segscan = nullptr;
                S.seed_cap = sized((size_t)n_seg) + 1024;
                CKD(cached_dev_malloc(&S.d_seeds, S.seed_cap * 8)); CKD(cached_dev_malloc(&S.d_segbase, S.seed_cap * 4));
                CKD(cached_dev_malloc(&S.d_segcnt, S.seed_cap * 4)); CKD(cached_dev_malloc(&S.d_segscan, S.seed_cap * 4));
            }
            if ((size_t)cap_total + 8 > S.rec_cap) {
                if (S.d_recoff) drain();
                const size_t cap = sized((size_t)cap_total) + 1024;
                cached_dev_free(S.d_recoff); cached_dev_free(S.d_dense); cached_dev_free(S.d_ncig); cached_dev_free(S.d_ncigscan); cached_dev_free(S.d_ws);
                S.d_recoff = S.d_dense = nullptr; S.d_ncig = S.d_ncigscan = nullptr; S.d_ws = nullptr;
                S.ws_cap = feed_scan_workspace_bytes((uint32_t)cap + 8);
                CKD(cached_dev_malloc(&S.d_recoff, cap * 4)); CKD(cached_dev_malloc(&S.d_dense, cap * 4)); CKD(cached_dev_malloc(&S.d_ncig, (cap + 8) * 4));
                CKD(cached_dev_malloc(&S.d_ncigscan, (cap + 8) * 4)); CKD(cached_dev_malloc(&S.d_ws, S.ws_cap));
                S.rec_cap = cap;
            }
            t_alloc += now_s() - ta0;
            uint8_t* data = S.d_infl + DeviceFeed::HEAD;
            const uint8_t* comp = resident ? F.d_file + G.first_coff : S.d_comp;
            cudaStream_t is = S.inf_stream, cs = stream_;
            // ---- inflate stream: block table, inflate, record starts
            // small tables go through the (pageable) vectors: cudaMemcpyAsync stages them synchronously, they are tiny
            CKD(cudaMemcpyAsync(S.d_desc, G.desc.data(), (size_t)nb * sizeof(BgzfBlockDesc), cudaMemcpyHostToDevice, is));
            CKD(cudaMemcpyAsync(S.d_seeds, sd.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, is));
            if (!resident) { CKD(cudaEventRecord(S.copied, copy_stream_)); CKD(cudaStreamWaitEvent(is, S.copied, 0)); }   // every staged chunk of the group has landed
            cudaEvent_t ea = nullptr, eb = nullptr;
            if (prm_.profile) { ea = get_event(); eb = get_event(); cudaEventRecord(ea, is); }
            launch_bgzf_inflate(comp, S.d_desc, nb, data, S.d_status, S.d_scratch, is);
            if (prm_.profile) { cudaEventRecord(eb, is); feed_prof_.push_back({ea, eb}); }
            if (seed_mode == 0) launch_block_seeds(data, (int64_t)G.out_total, S.d_desc, nb, n_ref, S.d_seeds, is);
            CKD(cudaEventRecord(S.inflated, is));                            // also: S.d_comp may be refilled
            // ---- chain stream: carry in, walk, gather, extract (append), carry out
            CKD(cudaMemcpyAsync(S.d_segbase, segbase.data(), ((size_t)n_seg + 1) * 4, cudaMemcpyHostToDevice, cs));
            CKD(cudaMemsetAsync(S.d_segcnt, 0, ((size_t)n_seg + 1) * 4, cs));
            launch_feed_reset(F.d_state, G.first_of_range ? 0 : 1, cs);
            CKD(cudaStreamWaitEvent(cs, S.inflated, 0));
            if (!G.first_of_range) launch_feed_carry_in(F.d_carry + DeviceFeed::HEAD, data, F.d_state, cs);
            launch_record_walk(data, (int64_t)G.out_total, limit, S.d_seeds, S.d_segbase, n_seg, G.first_of_range ? 0 : 1, F.d_state,
                               S.d_recoff, S.d_segcnt, cs);
            launch_record_gather(data, S.d_recoff, S.d_segbase, S.d_segcnt, S.d_segscan, n_seg, (uint32_t)cap_total, F.d_state, S.d_dense,
                                 S.d_ncig, S.d_ncigscan, S.d_ws, S.ws_cap, cs);
            launch_record_extract(data, S.d_dense, S.d_ncigscan, (uint32_t)cap_total, F.d_state, n_ref, xs_mode, (uint8_t)tag_[0], (uint8_t)tag_[1],
                                  F.a_tid, F.a_pos, F.a_meta, F.a_off, F.a_cigar, (uint32_t)std::min<size_t>(F.acc_rec_cap, 0xffffffffu),
                                  (uint32_t)std::min<size_t>(F.acc_ops_cap, 0xffffffffu), cs);
            launch_feed_finish(data, (int64_t)G.out_total, F.d_carry + DeviceFeed::HEAD, DeviceFeed::HEAD, S.d_segscan, n_seg, S.d_ncigscan, F.d_state, cs);
            launch_inflate_status_reduce(S.d_status, nb, &F.d_state->flags, cs);
            CKD(cudaMemcpyAsync(S.h_state, F.d_state, sizeof(FeedState), cudaMemcpyDeviceToHost, cs));
            CKD(cudaEventRecord(S.done, cs));
            CKD(cudaGetLastError());
            stats_.kernel_launches += 10 + (seed_mode == 0 ? 2 : 0) + (G.first_of_range ? 0 : 1);
            stats_.h2d_bytes += (size_t)nb * sizeof(BgzfBlockDesc) + (size_t)n_seg * 12;
            stats_.bgzf_blocks += nb; stats_.inflated_bytes += G.out_total;
            S.seq = ++launch_seq; S.cap_rec = cap_total; S.cap_ops = cig_upper;
            infl_rec += cap_total; infl_ops += cig_upper; acc_dirty = true;
            S.busy = true;
            S.dbg_nb = nb; S.dbg_nseg = n_seg; S.dbg_out_total = G.out_total; S.dbg_first = G.first_of_range;
            ++n_groups;
            return 0;
        };

        std::vector<BgzfBlockInfo> blocks;
        while (!stream_ends && !declined && !reached_limit) {
            DeviceFeed::GroupSlot& S = F.slot[n_groups % F.NSLOT];
            // ---- one chunk of compressed bytes: headers scanned on the host; page cache -> pinned -> the slot's device buffer
            blocks.clear();
            bool stop = false, partial = false, untrusted = false;
            const uint64_t c_first = coff;
            const uint8_t* src = nullptr;
            size_t got_bytes = 0;
            uint64_t scan_lim = end_coff;
            if (g.comp_bytes == 0) {
                // a new group starts: its slot must be free (the group that used it is checked here, NSLOT groups later)
                if ((rc = check(S))) return rc;
                if (declined || reached_limit) break;
            }
            if (resident) {
                src = bam.data() + c_first;
                got_bytes = (size_t)std::min<uint64_t>(STAGE + (1u << 16), bam.size() - c_first);
            } else {
                // windows sit on a fixed grid from the range's first block (that is what lets StagePipe read ahead); a chunk is the
                // blocks that START inside its window
                while (c_first >= win + STAGE) win += STAGE;
                NvtxRange nvtx("rtjx:stage window (wait for the read-ahead)");
                const double t0 = now_s();
                size_t win_got = 0;
                const uint8_t* w = pipe->get(win, end_coff, &win_got, &buf);
                t_stage += now_s() - t0; stats_.host_inflate_s += now_s() - t0;   // host staging time (nothing is inflated on the host)
                if (!w) return fail(RTJX_E_IO, "short read while staging the BAM");
                src = w + (c_first - win);
                got_bytes = win_got - (size_t)(c_first - win);
                scan_lim = std::min<uint64_t>(end_coff, win + STAGE);
            }
            const double ts0 = now_s();
            uint64_t c_end;
            if (resident) {
                auto it = F.scans.find(c_first);
                if (it == F.scans.end() || it->second.end_coff != end_coff || it->second.got != got_bytes) {
                    DeviceFeed::ChunkScan cs;
                    cs.c_end = scan_bgzf_blocks_mem(src, got_bytes, c_first, end_coff, &cs.blocks, &cs.stop, &cs.partial, &cs.untrusted);
                    cs.end_coff = end_coff; cs.got = got_bytes;
                    it = F.scans.insert_or_assign(c_first, std::move(cs)).first;
                }
                blocks = it->second.blocks; stop = it->second.stop; partial = it->second.partial; untrusted = it->second.untrusted;
                c_end = it->second.c_end;
            } else {
                c_end = scan_bgzf_blocks_mem(src, got_bytes, c_first, scan_lim, &blocks, &stop, &partial, &untrusted);
                if (stop && scan_lim < end_coff && c_end >= scan_lim) stop = false;   // the window ended, not the stream
            }
            t_scanhdr += now_s() - ts0;
            coff = c_end;
            if (untrusted) { declined = true; break; }     // a trailer that cannot be taken at its word: the host feeder decides
            if (stop || c_end >= end_coff || (partial && c_first + got_bytes >= bam.size()) || blocks.empty()) stream_ends = true;
            if (!blocks.empty()) {
                const size_t bytes = (size_t)(c_end - c_first);
                if (g.comp_bytes == 0) g.first_coff = c_first;
                if (!resident) {
                    if (g.comp_bytes == 0) {
                        CKD(cudaStreamWaitEvent(copy_stream_, S.inflated, 0));   // the inflate that last read this buffer
                        const size_t want = std::min<uint64_t>(GROUP + STAGE + (1u << 17), (end_coff - c_first) + (1u << 17)) + 256;
                        if (want > S.comp_cap) {
                            const double ta = now_s();
                            if (S.d_comp) drain();
                            cached_dev_free(S.d_comp); S.d_comp = nullptr;
                            CKD(cached_dev_malloc(&S.d_comp, want + want / 16));
                            S.comp_cap = want + want / 16;
                            t_alloc += now_s() - ta;
                        }
                    }
                    if (g.comp_bytes + bytes + 64 > S.comp_cap) return fail(RTJX_E_STATE, "device feed: group buffer overflow");
                    CKD(cudaMemcpyAsync(S.d_comp + g.comp_bytes, src, bytes, cudaMemcpyHostToDevice, copy_stream_));
                    CKD(cudaMemsetAsync(S.d_comp + g.comp_bytes + bytes, 0, 64, copy_stream_));   // the bit readers look a few words past a payload
                    CKD(cudaEventRecord(F.comp_free[buf], copy_stream_));
                    stats_.h2d_bytes += bytes;
                }
                for (const BgzfBlockInfo& b : blocks) {
                    BgzfBlockDesc d;
                    d.in_off = (uint32_t)(b.coff - g.first_coff) + 18; d.in_len = b.csize - 26;
                    d.out_off = (uint32_t)g.out_total; d.out_len = b.isize;
                    g.desc.push_back(d); g.coffs.push_back(b.coff); g.isize.push_back(b.isize);
                    g.out_total += b.isize;
                }
                g.comp_bytes += bytes;
                stats_.compressed_bytes += bytes;
            }
            // ---- close the group?
            // (by compressed bytes, or — a file that compresses better than ~3.5x — by inflated bytes: a group's stream must stay
            // well below 2 GiB)
            if (!g.desc.empty() && (g.comp_bytes >= (first_group ? std::min(GROUP, FIRST_GROUP) : GROUP) || g.out_total >= OUT_CAP || stream_ends)) {
                if ((rc = launch(g, S))) return rc;
                if (declined) break;
                first_group = false;
                g.clear(); g.first_of_range = false;
            }
        }
        // a range ends: everything it produced is scanned before the next range starts
        if ((rc = flush())) return rc;
    }
    drain();
    for (auto& pe : feed_prof_) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pe.first, pe.second) == cudaSuccess) stats_.inflate_kernel_ms += ms;
        ev_pool_.push_back(pe.first); ev_pool_.push_back(pe.second);
    }
    feed_prof_.clear();
    if (trace)
        fprintf(stderr, "[rtjx] device feed (seed mode %d, %s, %d groups): total %.1f ms | index seeds %.1f | header scan %.1f | staging memcpy %.1f | launch(host) %.1f (alloc %.1f) | slot wait %.1f | wait %.1f\n",
                seed_mode, resident ? "file resident in HBM" : "file staged through pinned memory", n_groups, 1e3 * (now_s() - t_begin), 1e3 * t_seeds,
                1e3 * t_scanhdr, 1e3 * t_stage, 1e3 * t_launch, 1e3 * t_alloc, 1e3 * t_slot, 1e3 * stats_.host_wait_s);
    if (declined) return 1;
    stats_.total_s += now_s() - t_begin;
    return RTJX_OK;
}

}  // namespace rtjx
