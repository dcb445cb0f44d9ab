// regtools_b200/csrc/device_run.cc — whole-file runs with BGZF inflate and BAM record split on the device.
//
// Host work per chunk of the file: walk BGZF block headers (18-byte header, BSIZE, ISIZE trailer —
// bgzf.c:348-355,525-546), memcpy the compressed bytes from the page cache into pinned memory with a
// few threads, translate the BAI's record-start virtual offsets that fall into the chunk into stream
// offsets (seeds), enqueue.  Everything else (inflate, record chain walk, field extraction, CIGAR scan,
// junction merge) runs on the GPU; the host reads back 40 bytes of counters per chunk.
// Any anomaly the device reports (malformed record, a walk that misses its next seed, capacity) makes
// Engine::run fall back to the host feeder for the whole run, so results never depend on this path
// accepting bad input.
#include "engine.h"
#include "buffer_cache.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
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
    if (n < (8u << 20) || threads <= 1) return one(dst, n, file_off);
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
    // double-buffered compressed staging
    uint8_t* h_comp[2] = {nullptr, nullptr};
    cudaEvent_t comp_free[2] = {nullptr, nullptr};
    cudaEvent_t inflate_done = nullptr, copies_done = nullptr;
    uint8_t* d_comp_group[2] = {nullptr, nullptr}; size_t group_cap[2] = {0, 0};   // compressed bytes of a group (double-buffered)
    cudaEvent_t group_free[2] = {nullptr, nullptr};                              // inflate of the group that used the buffer is done
    // per-chunk tables (host pinned + device)
    BgzfBlockDesc* d_desc = nullptr; size_t desc_cap = 0;
    int64_t* d_seeds = nullptr; uint32_t* d_segbase = nullptr; size_t seed_cap = 0;
    uint32_t* d_status = nullptr;
    void* d_inf_scratch = nullptr; size_t inf_scratch_cap = 0;   // match lists of the lane-per-stream inflate kernel
    uint8_t* d_infl = nullptr; size_t infl_cap = 0;      // HEAD + data + pad
    int32_t* d_recoff = nullptr; int32_t* d_dense = nullptr; uint32_t* d_ncig = nullptr; uint32_t* d_ncigscan = nullptr; size_t rec_cap = 0;
    uint32_t* d_segcnt = nullptr; uint32_t* d_segscan = nullptr;
    int32_t* d_tid = nullptr; int32_t* d_pos = nullptr; uint32_t* d_meta = nullptr; uint32_t* d_off = nullptr;
    uint32_t* d_cigar = nullptr; size_t cigar_cap = 0;
    void* d_ws = nullptr; size_t ws_cap = 0;
    FeedState* d_state = nullptr; FeedState* h_state = nullptr;
    uint32_t* h_status_sum = nullptr;

    ~DeviceFeed() {
        for (int i = 0; i < 2; ++i) {
            cached_host_free(h_comp[i]);
            if (comp_free[i]) cudaEventDestroy(comp_free[i]);
        }
        if (inflate_done) cudaEventDestroy(inflate_done);
        if (copies_done) cudaEventDestroy(copies_done);
        for (int i = 0; i < 2; ++i) { cached_dev_free(d_comp_group[i]); if (group_free[i]) cudaEventDestroy(group_free[i]); }
        cached_dev_free(d_desc); cached_dev_free(d_seeds); cached_dev_free(d_segbase); cached_dev_free(d_status); cached_dev_free(d_infl);
        cached_dev_free(d_recoff); cached_dev_free(d_dense); cached_dev_free(d_ncig); cached_dev_free(d_ncigscan); cached_dev_free(d_segcnt); cached_dev_free(d_segscan);
        cached_dev_free(d_tid); cached_dev_free(d_pos); cached_dev_free(d_meta); cached_dev_free(d_off); cached_dev_free(d_cigar); cached_dev_free(d_ws);
        cached_dev_free(d_state); cached_host_free(h_state); cached_dev_free(d_inf_scratch);
    }
};

void Engine::DeviceFeedDeleter::operator()(DeviceFeed* p) const { delete p; }

// One group = the unit of device work: a run of consecutive BGZF blocks (several staging chunks) that is
// inflated, split and scanned by one set of kernel launches.
struct FeedGroup {
    std::vector<BgzfBlockDesc> desc;     // in_off relative to the group's device buffer
    std::vector<uint64_t> coffs;         // compressed offset of every block (for seed lookup)
    std::vector<uint32_t> isize;
    uint64_t comp_bytes = 0, out_total = 0;
    uint64_t v_lo = 0;                   // first virtual offset of the group (record start for the first group of a range)
    bool first_of_range = false;
    void clear() { desc.clear(); coffs.clear(); isize.clear(); comp_bytes = out_total = 0; }
};

// Returns RTJX_OK, a negative status, or +1 = "device path declined, use the host feeder".
int Engine::run_device(const BamFile& bam, const BaiIndex& idx, const IterSpec& spec) {
    const double t_begin = now_s();
    int rc = ensure_device();
    if (rc) return rc;
    if (!dfeed_) dfeed_.reset(new DeviceFeed());
    DeviceFeed& F = *dfeed_;
    const int n_threads = prm_.n_threads > 0 ? prm_.n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
    const int copy_threads = std::min(n_threads, 8);
    const uint64_t STAGE = 16ull << 20;                  // compressed bytes per pinned staging chunk (pinned allocation costs ~0.5 ms/MB)
    static const uint64_t GROUP = [] { const char* v = getenv("RTJX_GROUP_MB"); return (uint64_t)(v ? atoi(v) : 512) << 20; }();
    // the first group of a range is smaller: the GPU starts after a few ms of staging instead of a full group's worth
    // (192 MB measured best on the 1.27 GB C2 file: 117 -> 107 ms end to end; irrelevant for files of many groups)
    static const uint64_t FIRST_GROUP = [] { const char* v = getenv("RTJX_FIRST_GROUP_MB"); return (uint64_t)(v ? atoi(v) : 192) << 20; }();

    // ---- ranges to stream (same as the host feeder)
    std::vector<Chunk64> ranges;
    if (spec.kind == IterSpec::WholeFile) {
        uint64_t off0;
        if (!idx.whole_file_start(&off0)) return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
        if (off0 == 0) off0 = bam.header().first_record_voffset;
        ranges.push_back(Chunk64{off0, UINT64_MAX});
    } else if (spec.kind == IterSpec::Contigs) {
        ranges = coalesced_contig_ranges(idx, spec.contigs);   // a shard of consecutive contigs streams as one range
    } else {
        return 1;
    }
    // ---- record-start virtual offsets known to the index
    std::vector<uint64_t> seeds_all;
    // (the 16 kb linear index: one record start per window that holds reads; already in file order for a sorted BAM)
    for (const BaiIndex::Ref& r : idx.refs)
        for (uint64_t v : r.ioffset) if (v && (seeds_all.empty() || v != seeds_all.back())) seeds_all.push_back(v);
    // (the binning index: every chunk begins at a record and ends right after one (hts_idx_push, hts.c:1288-1350).  Where
    // spliced and unspliced reads alternate between a leaf bin and its ancestors — exactly the deep-coverage windows whose
    // single 16 kb linear-index entry spans megabytes — the chunk list cuts the stream every few BGZF blocks: 5x the
    // seeds and half the longest segment on the C2 BAM.  The pseudo-bin's second "chunk" holds counts, not offsets.)
    if (!feed_linear_seeds_only_)
    for (const BaiIndex::Ref& r : idx.refs)
        for (const BaiIndex::Bin& b : r.bins) {
            if (b.bin == idx.meta_bin()) continue;
            for (const Chunk64& c : b.chunks) { if (c.beg) seeds_all.push_back(c.beg); if (c.end) seeds_all.push_back(c.end); }
        }
    if (!std::is_sorted(seeds_all.begin(), seeds_all.end())) std::sort(seeds_all.begin(), seeds_all.end());
    seeds_all.erase(std::unique(seeds_all.begin(), seeds_all.end()), seeds_all.end());

    // ---- fixed-size device state, pinned staging
    if (!F.d_state) {
        CKD(cached_dev_malloc(&F.d_state, sizeof(FeedState)));
        CKD(cached_host_alloc(&F.h_state, sizeof(FeedState)));
        for (int i = 0; i < 2; ++i) {
            CKD(cudaEventCreateWithFlags(&F.comp_free[i], cudaEventDisableTiming));
            CKD(cached_host_alloc(&F.h_comp[i], STAGE + (1u << 17) + 256));
        }
        CKD(cudaEventCreateWithFlags(&F.inflate_done, cudaEventDisableTiming));
        for (int i = 0; i < 2; ++i) CKD(cudaEventCreateWithFlags(&F.group_free[i], cudaEventDisableTiming));
        CKD(cudaEventCreateWithFlags(&F.copies_done, cudaEventDisableTiming));
    }
    auto grow_dev = [&](void** p, size_t* cap, size_t want, size_t elem) -> cudaError_t {
        if (want <= *cap) return cudaSuccess;
        if (*p) { cudaStreamSynchronize(stream_); cudaStreamSynchronize(copy_stream_); }   // only a live buffer needs the streams drained
        cached_dev_free(*p); *p = nullptr;
        size_t c = want + want / 8;
        cudaError_t e = cached_dev_malloc(p, c * elem + 256);
        if (e == cudaSuccess) *cap = c;
        return e;
    };

    static const bool trace = getenv("RTJX_TRACE") != nullptr;
    double t_stage = 0, t_launch = 0, t_scanhdr = 0, t_seeds = now_s() - t_begin, t_slot = 0, t_alloc = 0;
    const int xs_mode = prm_.strandness == 0;
    const int32_t n_ref = (int32_t)bam.header().names.size();
    uint64_t ordinal = 0;
    bool declined = false;
    bool in_flight = false;              // a group's kernels are enqueued and not yet harvested
    bool reached_limit = false;

    // harvest(): wait for the in-flight group, then launch cigar_scan + junction_merge on its batch
    auto harvest = [&]() -> int {
        if (!in_flight) return 0;
        in_flight = false;
        const double tw = now_s();
        CKD(cudaStreamSynchronize(stream_));
        stats_.host_wait_s += now_s() - tw;
        stats_.d2h_bytes += sizeof(FeedState);
        const FeedState st = *F.h_state;
        if (st.flags || st.bad_offset != LLONG_MAX) { declined = true; feed_decline_flags_ = st.flags; return 0; }
        if (st.n_rec) {
            BatchView v;
            v.n_reads = st.n_rec; v.n_ops = st.n_ops; v.first_ordinal = ordinal;
            v.tid = F.d_tid; v.pos = F.d_pos; v.meta = F.d_meta; v.cig_off = F.d_off; v.cigar = F.d_cigar;
            if (st.n_junction_ops) { int r = process_device_batch(v, st.n_junction_ops, stream_); if (r) return r; }
            else { stats_.reads += st.n_rec; stats_.cigar_ops += st.n_ops; stats_.batches++; }
            ordinal += st.n_rec;
        }
        if (st.reached_limit) reached_limit = true;
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
        int buf = 0, gbuf = 0;
        FeedGroup g;
        g.first_of_range = true; g.v_lo = rg.beg;

        // launch(g): tables H2D + inflate + walk + gather + extract + carry, all asynchronous on stream_
        auto launch = [&](FeedGroup& G) -> int {
            const double tl0 = now_s();
            struct TL { double* acc; double t0; ~TL() { *acc += now_s() - t0; } } tl{&t_launch, tl0};
            const uint32_t nb = (uint32_t)G.desc.size();
            // seeds: record starts inside this group, as stream offsets
            std::vector<int64_t> sd;
            sd.push_back(G.first_of_range ? (int64_t)(G.v_lo & 0xffff) : 0);
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
            const uint32_t n_seg = (uint32_t)sd.size();
            std::vector<uint32_t> segbase(n_seg + 1);
            uint64_t cap_total = 0;
            for (uint32_t i = 0; i < n_seg; ++i) {
                segbase[i] = (uint32_t)cap_total;
                const int64_t hi = i + 1 < n_seg ? sd[i + 1] : (int64_t)G.out_total;
                int64_t span = hi - sd[i];
                if (i == 0 && !G.first_of_range) span += DeviceFeed::HEAD;          // segment 0 starts inside the carry
                cap_total += (uint64_t)(span / 64 + 4);     // records are >= 37 bytes; denser than 64 B/record -> capacity flag -> host path
            }
            segbase[n_seg] = (uint32_t)cap_total;
            if (cap_total > 0x7ffffff0ull || G.out_total > 0x7ff00000ull) return fail(RTJX_E_STATE, "device feed: group too large");
            // range end inside this group?
            int64_t limit = LLONG_MAX;
            if (bounded) {
                const uint64_t ec = rg.end >> 16, eu = rg.end & 0xffff;
                size_t k = std::lower_bound(G.coffs.begin(), G.coffs.end(), ec) - G.coffs.begin();
                if (k < nb && G.coffs[k] == ec) limit = (int64_t)G.desc[k].out_off + (int64_t)eu;
                else if (k == nb && eu == 0 && G.coffs.back() < ec) limit = (int64_t)G.out_total;
            }
            // device buffers
            const double ta0 = now_s();
            struct TA2 { double* acc; double t0; bool on; ~TA2() { if (on) *acc += now_s() - t0; } } ta2{&t_alloc, ta0, true};
            if (DeviceFeed::HEAD + G.out_total + 64 > F.infl_cap) {
                // the headroom of the old buffer holds the record carried over from the previous group: keep it
                if (F.d_infl) cudaStreamSynchronize(stream_);
                const size_t want = DeviceFeed::HEAD + G.out_total + 64;
                const size_t cap = want + want / 8;
                uint8_t* nb2 = nullptr;
                CKD(cached_dev_malloc(&nb2, cap + 256));
                if (F.d_infl) CKD(cudaMemcpy(nb2, F.d_infl, DeviceFeed::HEAD, cudaMemcpyDeviceToDevice));
                cached_dev_free(F.d_infl);
                F.d_infl = nb2; F.infl_cap = cap;
            }
            CKD(grow_dev((void**)&F.d_cigar, &F.cigar_cap, (DeviceFeed::HEAD + G.out_total) / 32 + 4096, 4));   // > 12.5 % of the bytes being CIGAR -> capacity flag
            if ((size_t)nb + 2 > F.desc_cap) {
                if (F.d_desc) cudaStreamSynchronize(stream_);
                cached_dev_free(F.d_desc); cached_dev_free(F.d_status); F.d_desc = nullptr; F.d_status = nullptr;
                F.desc_cap = (size_t)nb * 2 + 1024;
                CKD(cached_dev_malloc(&F.d_desc, F.desc_cap * sizeof(BgzfBlockDesc))); CKD(cached_dev_malloc(&F.d_status, F.desc_cap * 4));
            }
            CKD(grow_dev(&F.d_inf_scratch, &F.inf_scratch_cap, bgzf_inflate_scratch_bytes(nb), 1));
            if ((size_t)n_seg + 2 > F.seed_cap) {
                if (F.d_seeds) cudaStreamSynchronize(stream_);
                cached_dev_free(F.d_seeds); cached_dev_free(F.d_segbase); cached_dev_free(F.d_segcnt); cached_dev_free(F.d_segscan);
                F.d_seeds = nullptr; F.d_segbase = F.d_segcnt = F.d_segscan = nullptr;
                F.seed_cap = (size_t)n_seg * 2 + 1024;
                CKD(cached_dev_malloc(&F.d_seeds, F.seed_cap * 8)); CKD(cached_dev_malloc(&F.d_segbase, F.seed_cap * 4));
                CKD(cached_dev_malloc(&F.d_segcnt, F.seed_cap * 4)); CKD(cached_dev_malloc(&F.d_segscan, F.seed_cap * 4));
            }
            if ((size_t)cap_total + 8 > F.rec_cap) {
                if (F.d_recoff) cudaStreamSynchronize(stream_);
                const size_t cap = (size_t)cap_total + cap_total / 8 + 1024;
                cached_dev_free(F.d_recoff); cached_dev_free(F.d_dense); cached_dev_free(F.d_ncig); cached_dev_free(F.d_ncigscan);
                cached_dev_free(F.d_tid); cached_dev_free(F.d_pos); cached_dev_free(F.d_meta); cached_dev_free(F.d_off); cached_dev_free(F.d_ws);
                F.d_recoff = F.d_dense = nullptr; F.d_ncig = F.d_ncigscan = nullptr; F.d_tid = F.d_pos = nullptr; F.d_meta = F.d_off = nullptr; F.d_ws = nullptr;
                F.ws_cap = feed_scan_workspace_bytes((uint32_t)cap + 8);
                CKD(cached_dev_malloc(&F.d_recoff, cap * 4)); CKD(cached_dev_malloc(&F.d_dense, cap * 4)); CKD(cached_dev_malloc(&F.d_ncig, (cap + 8) * 4));
                CKD(cached_dev_malloc(&F.d_ncigscan, (cap + 8) * 4)); CKD(cached_dev_malloc(&F.d_tid, cap * 4)); CKD(cached_dev_malloc(&F.d_pos, cap * 4));
                CKD(cached_dev_malloc(&F.d_meta, cap * 4)); CKD(cached_dev_malloc(&F.d_off, (cap + 8) * 4)); CKD(cached_dev_malloc(&F.d_ws, F.ws_cap));
                F.rec_cap = cap;
            }
            t_alloc += now_s() - ta0; ta2.on = false;
            uint8_t* data = F.d_infl + DeviceFeed::HEAD;
            // small tables go through the (pageable) vectors: cudaMemcpyAsync stages them synchronously, they are tiny
            CKD(cudaMemcpyAsync(F.d_desc, G.desc.data(), (size_t)nb * sizeof(BgzfBlockDesc), cudaMemcpyHostToDevice, stream_));
            CKD(cudaMemcpyAsync(F.d_seeds, sd.data(), (size_t)n_seg * 8, cudaMemcpyHostToDevice, stream_));
            CKD(cudaMemcpyAsync(F.d_segbase, segbase.data(), ((size_t)n_seg + 1) * 4, cudaMemcpyHostToDevice, stream_));
            CKD(cudaMemsetAsync(F.d_segcnt, 0, ((size_t)n_seg + 1) * 4, stream_));
            launch_feed_reset(F.d_state, G.first_of_range ? 0 : 1, stream_);
            CKD(cudaEventRecord(F.copies_done, copy_stream_));
            CKD(cudaStreamWaitEvent(stream_, F.copies_done, 0));             // every staged chunk of the group has landed
            cudaEvent_t ea = nullptr, eb = nullptr;
            if (prm_.profile) { ea = get_event(); eb = get_event(); cudaEventRecord(ea, stream_); }
            launch_bgzf_inflate(F.d_comp_group[gbuf], F.d_desc, nb, data, F.d_status, F.d_inf_scratch, stream_);
            if (prm_.profile) { cudaEventRecord(eb, stream_); feed_prof_.push_back({ea, eb}); }
            CKD(cudaEventRecord(F.group_free[gbuf], stream_));               // this compressed buffer may be refilled after the inflate
            launch_record_walk(data, (int64_t)G.out_total, limit, F.d_seeds, F.d_segbase, n_seg, G.first_of_range ? 0 : 1, F.d_state,
                               F.d_recoff, F.d_segcnt, stream_);
            launch_record_gather(data, F.d_recoff, F.d_segbase, F.d_segcnt, F.d_segscan, n_seg, (uint32_t)cap_total, F.d_state, F.d_dense,
                                 F.d_ncig, F.d_ncigscan, F.d_ws, F.ws_cap, stream_);
            launch_record_extract(data, F.d_dense, F.d_ncigscan, (uint32_t)cap_total, F.d_state, n_ref, xs_mode, (uint8_t)tag_[0], (uint8_t)tag_[1],
                                  F.d_tid, F.d_pos, F.d_meta, F.d_off, F.d_cigar, (uint32_t)std::min<size_t>(F.cigar_cap, 0xffffffffu), stream_);
            launch_feed_finish(data, (int64_t)G.out_total, data, DeviceFeed::HEAD, F.d_segscan, n_seg, F.d_ncigscan, F.d_state, stream_);
            launch_inflate_status_reduce(F.d_status, nb, &F.d_state->flags, stream_);
            CKD(cudaMemcpyAsync(F.h_state, F.d_state, sizeof(FeedState), cudaMemcpyDeviceToHost, stream_));
            CKD(cudaGetLastError());
            stats_.kernel_launches += 9;     // reset, inflate, walk, count, gather, ops, extract, finish, status (+2 CUB scans)
            stats_.h2d_bytes += (size_t)nb * sizeof(BgzfBlockDesc) + (size_t)n_seg * 12;
            stats_.bgzf_blocks += nb; stats_.inflated_bytes += G.out_total;
            in_flight = true;
            return 0;
        };

        std::vector<BgzfBlockInfo> blocks;
        while (!stream_ends && !declined && !reached_limit) {
            // ---- stage one chunk of compressed bytes: page cache -> pinned -> device (group buffer)
            blocks.clear();
            bool stop = false, partial = false, untrusted = false;
            const uint64_t c_first = coff;
            { const double tq = now_s(); CKD(cudaEventSynchronize(F.comp_free[buf])); t_slot += now_s() - tq; }   // pinned slot free (its H2D two chunks back is done)
            const size_t want_bytes = (size_t)std::min<uint64_t>(STAGE + (1u << 16), bam.size() - c_first);
            const double t0 = now_s();
            const size_t got_bytes = parallel_pread(bam.fd(), F.h_comp[buf], want_bytes, c_first, copy_threads);
            t_stage += now_s() - t0; stats_.host_inflate_s += now_s() - t0;   // host staging time (nothing is inflated on the host)
            const double ts0 = now_s();
            const uint64_t c_end = scan_bgzf_blocks_mem(F.h_comp[buf], got_bytes, c_first, end_coff, &blocks, &stop, &partial, &untrusted);
            t_scanhdr += now_s() - ts0;
            coff = c_end;
            if (untrusted) { declined = true; break; }     // a trailer that cannot be taken at its word: the host feeder decides
            if (stop || c_end >= end_coff || (partial && c_first + got_bytes >= bam.size()) || blocks.empty()) stream_ends = true;
            if (!blocks.empty()) {
                const size_t bytes = (size_t)(c_end - c_first);
                // the group buffer must hold this chunk; it may still be read by the previous group's inflate
                if (g.comp_bytes == 0) {
                    CKD(cudaStreamWaitEvent(copy_stream_, F.group_free[gbuf], 0));   // the inflate that last read this buffer
                    const size_t want = std::min<uint64_t>(GROUP + STAGE + (1u << 17), (end_coff - c_first) + (1u << 17)) + 256;
                    if (want > F.group_cap[gbuf]) {
                        const double ta = now_s();
                        struct TA { double* acc; double t0; ~TA() { *acc += now_s() - t0; } } ta_guard{&t_alloc, ta};
                        if (F.d_comp_group[gbuf]) { cudaStreamSynchronize(stream_); cudaStreamSynchronize(copy_stream_); }
                        cached_dev_free(F.d_comp_group[gbuf]); F.d_comp_group[gbuf] = nullptr;
                        CKD(cached_dev_malloc(&F.d_comp_group[gbuf], want + want / 16));
                        F.group_cap[gbuf] = want + want / 16;
                    }
                }
                if (g.comp_bytes + bytes + 64 > F.group_cap[gbuf]) return fail(RTJX_E_STATE, "device feed: group buffer overflow");
                memset(F.h_comp[buf] + bytes, 0, 64);
                CKD(cudaMemcpyAsync(F.d_comp_group[gbuf] + g.comp_bytes, F.h_comp[buf], bytes + 64, cudaMemcpyHostToDevice, copy_stream_));
                CKD(cudaEventRecord(F.comp_free[buf], copy_stream_));
                buf ^= 1;
                for (const BgzfBlockInfo& b : blocks) {
                    BgzfBlockDesc d;
                    d.in_off = (uint32_t)(g.comp_bytes + (b.coff - c_first)) + 18; d.in_len = b.csize - 26;
                    d.out_off = (uint32_t)g.out_total; d.out_len = b.isize;
                    g.desc.push_back(d); g.coffs.push_back(b.coff); g.isize.push_back(b.isize);
                    g.out_total += b.isize;
                }
                g.comp_bytes += bytes;
                stats_.h2d_bytes += bytes; stats_.compressed_bytes += bytes;
            }
            // ---- close the group?
            if (!g.desc.empty() && (g.comp_bytes >= (first_group ? std::min(GROUP, FIRST_GROUP) : GROUP) || stream_ends)) {
                if ((rc = harvest())) return rc;                              // previous group: scan + merge (its SoA is about to be overwritten)
                if (declined || reached_limit) break;
                if ((rc = launch(g))) return rc;
                first_group = false;
                g.clear(); g.first_of_range = false;
                gbuf ^= 1;
            }
        }
        if ((rc = harvest())) return rc;
        (void)first_group;
    }
    CKD(cudaStreamSynchronize(copy_stream_));
    CKD(cudaStreamSynchronize(stream_));
    for (auto& pe : feed_prof_) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pe.first, pe.second) == cudaSuccess) stats_.inflate_kernel_ms += ms;
        ev_pool_.push_back(pe.first); ev_pool_.push_back(pe.second);
    }
    feed_prof_.clear();
    if (trace)
        fprintf(stderr, "[rtjx] device feed: total %.1f ms | index seeds %.1f | header scan %.1f | staging memcpy %.1f | launch(host) %.1f (alloc %.1f) | slot wait %.1f | wait %.1f\n",
                1e3 * (now_s() - t_begin), 1e3 * t_seeds, 1e3 * t_scanhdr, 1e3 * t_stage, 1e3 * t_launch, 1e3 * t_alloc, 1e3 * t_slot, 1e3 * stats_.host_wait_s);
    if (declined) return 1;
    stats_.total_s += now_s() - t_begin;
    return RTJX_OK;
}

}  // namespace rtjx
