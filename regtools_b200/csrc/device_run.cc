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

#include <algorithm>
#include <chrono>
#include <climits>
#include <cstring>
#include <thread>

namespace rtjx {

namespace {
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void parallel_memcpy(uint8_t* dst, const uint8_t* src, size_t n, int threads) {
    if (n < (8u << 20) || threads <= 1) { memcpy(dst, src, n); return; }
    std::vector<std::thread> pool;
    const size_t per = ((n + threads - 1) / threads + 4095) & ~(size_t)4095;
    for (int t = 0; t < threads; ++t) {
        const size_t o = (size_t)t * per;
        if (o >= n) break;
        const size_t len = std::min(per, n - o);
        pool.emplace_back([=] { memcpy(dst + o, src + o, len); });
    }
    for (auto& th : pool) th.join();
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
    uint8_t* h_comp[2] = {nullptr, nullptr}; uint8_t* d_comp[2] = {nullptr, nullptr}; size_t comp_cap = 0;
    cudaEvent_t comp_free[2] = {nullptr, nullptr};
    // per-chunk tables (host pinned + device)
    BgzfBlockDesc* h_desc[2] = {nullptr, nullptr}; BgzfBlockDesc* d_desc = nullptr; size_t desc_cap = 0;
    int64_t* h_seeds[2] = {nullptr, nullptr}; int64_t* d_seeds = nullptr; uint32_t* h_segbase[2] = {nullptr, nullptr};
    uint32_t* d_segbase = nullptr; size_t seed_cap = 0;
    uint32_t* d_status = nullptr;
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
            cudaFreeHost(h_comp[i]); cudaFree(d_comp[i]); cudaFreeHost(h_desc[i]); cudaFreeHost(h_seeds[i]); cudaFreeHost(h_segbase[i]);
            if (comp_free[i]) cudaEventDestroy(comp_free[i]);
        }
        cudaFree(d_desc); cudaFree(d_seeds); cudaFree(d_segbase); cudaFree(d_status); cudaFree(d_infl);
        cudaFree(d_recoff); cudaFree(d_dense); cudaFree(d_ncig); cudaFree(d_ncigscan); cudaFree(d_segcnt); cudaFree(d_segscan);
        cudaFree(d_tid); cudaFree(d_pos); cudaFree(d_meta); cudaFree(d_off); cudaFree(d_cigar); cudaFree(d_ws);
        cudaFree(d_state); cudaFreeHost(h_state);
    }
};

void Engine::DeviceFeedDeleter::operator()(DeviceFeed* p) const { delete p; }

// One chunk as prepared by the host.
struct FeedChunk {
    std::vector<BgzfBlockInfo> blocks;
    uint64_t coff_first = 0, coff_end = 0;
    uint64_t out_total = 0;
    uint32_t n_seg = 0, cap_total = 0;
    int64_t limit = LLONG_MAX;
    bool first_of_range = false, stream_ends = false;
    int buf = 0;
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
    const uint64_t CHUNK = 96ull << 20;                  // compressed bytes per chunk

    // ---- ranges to stream (same as the host feeder)
    std::vector<Chunk64> ranges;
    if (spec.kind == IterSpec::WholeFile) {
        uint64_t off0;
        if (!idx.whole_file_start(&off0)) return fail(RTJX_E_REGION, "Unable to iterate to region within BAM.\n\n");
        if (off0 == 0) off0 = bam.header().first_record_voffset;
        ranges.push_back(Chunk64{off0, UINT64_MAX});
    } else if (spec.kind == IterSpec::Contigs) {
        for (int32_t tid : spec.contigs) { Chunk64 c; if (idx.contig_range(tid, &c)) ranges.push_back(c); }
    } else {
        return 1;
    }
    // ---- record-start virtual offsets known to the index
    std::vector<uint64_t> seeds_all;
    for (const BaiIndex::Ref& r : idx.refs) {
        for (uint64_t v : r.ioffset) if (v) seeds_all.push_back(v);
        for (const BaiIndex::Bin& b : r.bins) {
            if (b.bin >= 37449u) continue;
            for (const Chunk64& c : b.chunks) seeds_all.push_back(c.beg);
        }
    }
    std::sort(seeds_all.begin(), seeds_all.end());
    seeds_all.erase(std::unique(seeds_all.begin(), seeds_all.end()), seeds_all.end());

    // ---- fixed-size device state
    if (!F.d_state) {
        CKD(cudaMalloc(&F.d_state, sizeof(FeedState)));
        CKD(cudaHostAlloc(&F.h_state, sizeof(FeedState), cudaHostAllocDefault));
        for (int i = 0; i < 2; ++i) CKD(cudaEventCreateWithFlags(&F.comp_free[i], cudaEventDisableTiming));
    }
    auto grow_dev = [&](void** p, size_t* cap, size_t want, size_t elem) -> cudaError_t {
        if (want <= *cap) return cudaSuccess;
        cudaStreamSynchronize(stream_);
        cudaFree(*p); *p = nullptr;
        size_t c = want + want / 4;
        cudaError_t e = cudaMalloc(p, c * elem + 64);
        if (e == cudaSuccess) *cap = c;
        return e;
    };

    const FeederOptions fo_dummy;
    (void)fo_dummy;
    const int xs_mode = prm_.strandness == 0;
    const int32_t n_ref = (int32_t)bam.header().names.size();
    uint64_t ordinal = 0;
    bool declined = false;

    for (size_t ri = 0; ri < ranges.size() && !declined; ++ri) {
        const Chunk64 rg = ranges[ri];
        uint64_t coff = rg.beg >> 16;
        const bool bounded = rg.end != UINT64_MAX;
        // blocks needed: up to the one holding the end offset (exclusive if the end sits on a block boundary)
        const uint64_t end_coff = bounded ? ((rg.end & 0xffff) ? (rg.end >> 16) + 1 : (rg.end >> 16)) : bam.size();
        bool first = true, range_done = false;
        int buf = 0;

        // prepare(k): scan headers, stage compressed bytes, build seeds  (host only)
        auto prepare = [&](FeedChunk& c) -> int {
            c.blocks.clear(); c.buf = buf; buf ^= 1;
            bool stop = false;
            c.coff_first = coff;
            c.coff_end = scan_bgzf_blocks(bam, coff, std::min<uint64_t>(end_coff, bam.size()), (size_t)-1, CHUNK, &c.blocks, &stop);
            c.stream_ends = stop;
            coff = c.coff_end;
            if (c.blocks.empty()) return 0;
            const size_t comp_bytes = (size_t)(c.coff_end - c.coff_first);
            if (comp_bytes + 64 > F.comp_cap) {
                cudaStreamSynchronize(stream_);
                const size_t cap = std::max<size_t>(comp_bytes + comp_bytes / 8, (size_t)CHUNK + (1u << 20)) + 64;
                for (int i = 0; i < 2; ++i) {
                    cudaFreeHost(F.h_comp[i]); cudaFree(F.d_comp[i]); F.h_comp[i] = F.d_comp[i] = nullptr;
                    if (cudaHostAlloc(&F.h_comp[i], cap, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&F.d_comp[i], cap) != cudaSuccess)
                        return fail(RTJX_E_CUDA, "device feed: staging allocation failed");
                }
                F.comp_cap = cap;
            }
            if (c.blocks.size() + 2 > F.desc_cap) {
                cudaStreamSynchronize(stream_);
                const size_t cap = c.blocks.size() * 2 + 1024;
                for (int i = 0; i < 2; ++i) { cudaFreeHost(F.h_desc[i]); F.h_desc[i] = nullptr; if (cudaHostAlloc(&F.h_desc[i], cap * sizeof(BgzfBlockDesc), cudaHostAllocDefault) != cudaSuccess) return fail(RTJX_E_CUDA, "device feed: allocation failed"); }
                cudaFree(F.d_desc); cudaFree(F.d_status); F.d_desc = nullptr; F.d_status = nullptr;
                if (cudaMalloc(&F.d_desc, cap * sizeof(BgzfBlockDesc)) != cudaSuccess || cudaMalloc(&F.d_status, cap * 4) != cudaSuccess) return fail(RTJX_E_CUDA, "device feed: allocation failed");
                F.desc_cap = cap;
            }
            // the staging buffer of this slot may still be read by an H2D copy two chunks back
            cudaEventSynchronize(F.comp_free[c.buf]);
            const double t0 = now_s();
            parallel_memcpy(F.h_comp[c.buf], bam.data() + c.coff_first, comp_bytes, copy_threads);
            memset(F.h_comp[c.buf] + comp_bytes, 0, 32);
            stats_.host_inflate_s += now_s() - t0;              // host staging time (no inflate happens on the host)
            uint64_t out = 0;
            BgzfBlockDesc* d = F.h_desc[c.buf];
            for (size_t i = 0; i < c.blocks.size(); ++i) {
                d[i].in_off = (uint32_t)(c.blocks[i].coff - c.coff_first) + 18; d[i].in_len = c.blocks[i].csize - 26;
                d[i].out_off = (uint32_t)out; d[i].out_len = c.blocks[i].isize;
                out += c.blocks[i].isize;
            }
            c.out_total = out;
            // seeds: record starts inside this chunk, as stream offsets
            std::vector<int64_t> sd;
            c.first_of_range = first;
            const uint64_t v_lo = first ? rg.beg : (c.coff_first << 16);
            const uint64_t v_hi = c.coff_end << 16;
            sd.push_back(first ? (int64_t)(rg.beg & 0xffff) : 0);
            auto it = std::upper_bound(seeds_all.begin(), seeds_all.end(), v_lo);
            size_t bi = 0;
            for (; it != seeds_all.end() && *it < v_hi; ++it) {
                if (bounded && *it >= rg.end) break;
                const uint64_t sc = *it >> 16, su = *it & 0xffff;
                while (bi < c.blocks.size() && c.blocks[bi].coff < sc) ++bi;
                if (bi == c.blocks.size() || c.blocks[bi].coff != sc || su > c.blocks[bi].isize) continue;   // not a block of this file
                const int64_t o = (int64_t)d[bi].out_off + (int64_t)su;
                if (o > sd.back()) sd.push_back(o);
            }
            c.n_seg = (uint32_t)sd.size();
            if (sd.size() + 2 > F.seed_cap) {
                cudaStreamSynchronize(stream_);
                const size_t cap = sd.size() * 2 + 1024;
                for (int i = 0; i < 2; ++i) {
                    cudaFreeHost(F.h_seeds[i]); cudaFreeHost(F.h_segbase[i]); F.h_seeds[i] = nullptr; F.h_segbase[i] = nullptr;
                    if (cudaHostAlloc(&F.h_seeds[i], cap * 8, cudaHostAllocDefault) != cudaSuccess || cudaHostAlloc(&F.h_segbase[i], cap * 4, cudaHostAllocDefault) != cudaSuccess)
                        return fail(RTJX_E_CUDA, "device feed: allocation failed");
                }
                cudaFree(F.d_seeds); cudaFree(F.d_segbase); cudaFree(F.d_segcnt); cudaFree(F.d_segscan);
                F.d_seeds = nullptr; F.d_segbase = F.d_segcnt = F.d_segscan = nullptr;
                if (cudaMalloc(&F.d_seeds, cap * 8) != cudaSuccess || cudaMalloc(&F.d_segbase, cap * 4) != cudaSuccess ||
                    cudaMalloc(&F.d_segcnt, cap * 4) != cudaSuccess || cudaMalloc(&F.d_segscan, cap * 4) != cudaSuccess)
                    return fail(RTJX_E_CUDA, "device feed: allocation failed");
                F.seed_cap = cap;
            }
            uint64_t cap_total = 0;
            for (size_t i = 0; i < sd.size(); ++i) {
                F.h_seeds[c.buf][i] = sd[i];
                F.h_segbase[c.buf][i] = (uint32_t)cap_total;
                const int64_t hi = i + 1 < sd.size() ? sd[i + 1] : (int64_t)out;
                int64_t span = hi - sd[i];
                if (i == 0 && !first) span += DeviceFeed::HEAD;          // segment 0 starts inside the carry
                cap_total += (uint64_t)(span / 36 + 2);
            }
            F.h_segbase[c.buf][sd.size()] = (uint32_t)cap_total;
            c.cap_total = (uint32_t)cap_total;
            // range end inside this chunk?
            c.limit = LLONG_MAX;
            if (bounded) {
                const uint64_t ec = rg.end >> 16, eu = rg.end & 0xffff;
                if (ec >= c.coff_first && ec < c.coff_end) {
                    size_t k = 0;
                    while (k < c.blocks.size() && c.blocks[k].coff < ec) ++k;
                    if (k < c.blocks.size() && c.blocks[k].coff == ec) c.limit = (int64_t)d[k].out_off + (int64_t)eu;
                } else if (ec == c.coff_end && eu == 0) c.limit = (int64_t)out;
            }
            first = false;
            return 0;
        };

        // enqueue(k): H2D + inflate + walk + gather + extract + carry  (all asynchronous on stream_)
        auto enqueue = [&](const FeedChunk& c) -> int {
            const size_t comp_bytes = (size_t)(c.coff_end - c.coff_first) + 32;
            CKD(grow_dev((void**)&F.d_infl, &F.infl_cap, DeviceFeed::HEAD + c.out_total + 64, 1));
            CKD(grow_dev((void**)&F.d_cigar, &F.cigar_cap, (DeviceFeed::HEAD + c.out_total) / 4 + 64, 4));
            if ((size_t)c.cap_total + 8 > F.rec_cap) {
                cudaStreamSynchronize(stream_);
                const size_t cap = (size_t)c.cap_total + c.cap_total / 4 + 1024;
                cudaFree(F.d_recoff); cudaFree(F.d_dense); cudaFree(F.d_ncig); cudaFree(F.d_ncigscan);
                cudaFree(F.d_tid); cudaFree(F.d_pos); cudaFree(F.d_meta); cudaFree(F.d_off); cudaFree(F.d_ws);
                F.d_recoff = F.d_dense = nullptr; F.d_ncig = F.d_ncigscan = nullptr; F.d_tid = F.d_pos = nullptr; F.d_meta = F.d_off = nullptr; F.d_ws = nullptr;
                F.ws_cap = feed_scan_workspace_bytes((uint32_t)cap + 8);
                CKD(cudaMalloc(&F.d_recoff, cap * 4)); CKD(cudaMalloc(&F.d_dense, cap * 4)); CKD(cudaMalloc(&F.d_ncig, (cap + 8) * 4));
                CKD(cudaMalloc(&F.d_ncigscan, (cap + 8) * 4)); CKD(cudaMalloc(&F.d_tid, cap * 4)); CKD(cudaMalloc(&F.d_pos, cap * 4));
                CKD(cudaMalloc(&F.d_meta, cap * 4)); CKD(cudaMalloc(&F.d_off, (cap + 8) * 4)); CKD(cudaMalloc(&F.d_ws, F.ws_cap));
                F.rec_cap = cap;
            }
            uint8_t* data = F.d_infl + DeviceFeed::HEAD;
            CKD(cudaMemcpyAsync(F.d_comp[c.buf], F.h_comp[c.buf], comp_bytes, cudaMemcpyHostToDevice, stream_));
            CKD(cudaEventRecord(F.comp_free[c.buf], stream_));
            CKD(cudaMemcpyAsync(F.d_desc, F.h_desc[c.buf], c.blocks.size() * sizeof(BgzfBlockDesc), cudaMemcpyHostToDevice, stream_));
            CKD(cudaMemcpyAsync(F.d_seeds, F.h_seeds[c.buf], (size_t)c.n_seg * 8, cudaMemcpyHostToDevice, stream_));
            CKD(cudaMemcpyAsync(F.d_segbase, F.h_segbase[c.buf], ((size_t)c.n_seg + 1) * 4, cudaMemcpyHostToDevice, stream_));
            CKD(cudaMemsetAsync(F.d_segcnt, 0, ((size_t)c.n_seg + 1) * 4, stream_));
            launch_feed_reset(F.d_state, c.first_of_range ? 0 : 1, stream_);
            cudaEvent_t ea = nullptr, eb = nullptr;
            if (prm_.profile) { ea = get_event(); eb = get_event(); cudaEventRecord(ea, stream_); }
            launch_bgzf_inflate(F.d_comp[c.buf], F.d_desc, (uint32_t)c.blocks.size(), data, F.d_status, stream_);
            if (prm_.profile) { cudaEventRecord(eb, stream_); feed_prof_.push_back({ea, eb}); }
            launch_record_walk(data, (int64_t)c.out_total, c.limit, F.d_seeds, F.d_segbase, c.n_seg, c.first_of_range ? 0 : 1, F.d_state,
                               F.d_recoff, F.d_segcnt, stream_);
            launch_record_gather(data, F.d_recoff, F.d_segbase, F.d_segcnt, F.d_segscan, c.n_seg, c.cap_total, F.d_state, F.d_dense,
                                 F.d_ncig, F.d_ncigscan, F.d_ws, F.ws_cap, stream_);
            launch_record_extract(data, F.d_dense, F.d_ncigscan, c.cap_total, F.d_state, n_ref, xs_mode, (uint8_t)tag_[0], (uint8_t)tag_[1],
                                  F.d_tid, F.d_pos, F.d_meta, F.d_off, F.d_cigar, stream_);
            launch_feed_finish(data, (int64_t)c.out_total, data, DeviceFeed::HEAD, F.d_segscan, c.n_seg, F.d_ncigscan, F.d_state, stream_);
            launch_inflate_status_reduce(F.d_status, (uint32_t)c.blocks.size(), &F.d_state->flags, stream_);
            CKD(cudaMemcpyAsync(F.h_state, F.d_state, sizeof(FeedState), cudaMemcpyDeviceToHost, stream_));
            CKD(cudaGetLastError());
            stats_.kernel_launches += 9;     // reset, inflate, walk, count, gather, ops, extract, finish, status (+2 CUB scans)
            stats_.h2d_bytes += comp_bytes + c.blocks.size() * sizeof(BgzfBlockDesc) + (size_t)c.n_seg * 12;
            stats_.bgzf_blocks += c.blocks.size(); stats_.compressed_bytes += comp_bytes - 32; stats_.inflated_bytes += c.out_total;
            return 0;
        };

        FeedChunk cur, nxt;
        if ((rc = prepare(cur))) return rc;
        if (cur.blocks.empty()) continue;
        if ((rc = enqueue(cur))) return rc;
        while (!range_done && !declined) {
            const bool more = !cur.stream_ends && cur.coff_end < std::min<uint64_t>(end_coff, bam.size());
            if (more) { if ((rc = prepare(nxt))) return rc; }          // host staging overlaps the GPU work of `cur`
            const double tw = now_s();
            CKD(cudaStreamSynchronize(stream_));
            stats_.host_wait_s += now_s() - tw;
            stats_.d2h_bytes += sizeof(FeedState);
            const FeedState st = *F.h_state;
            if (st.flags || st.bad_offset != LLONG_MAX) { declined = true; break; }
            if (st.n_rec) {
                BatchView v;
                v.n_reads = st.n_rec; v.n_ops = st.n_ops; v.first_ordinal = ordinal;
                v.tid = F.d_tid; v.pos = F.d_pos; v.meta = F.d_meta; v.cig_off = F.d_off; v.cigar = F.d_cigar;
                if (st.n_junction_ops) { if ((rc = process_device_batch(v, st.n_junction_ops, stream_))) return rc; }
                else { stats_.reads += st.n_rec; stats_.cigar_ops += st.n_ops; stats_.batches++; }
                ordinal += st.n_rec;
            }
            if (st.reached_limit || !more || nxt.blocks.empty()) { range_done = true; break; }
            if ((rc = enqueue(nxt))) return rc;
            std::swap(cur, nxt);
        }
    }
    CKD(cudaStreamSynchronize(stream_));
    for (auto& pe : feed_prof_) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, pe.first, pe.second) == cudaSuccess) stats_.inflate_kernel_ms += ms;
        ev_pool_.push_back(pe.first); ev_pool_.push_back(pe.second);
    }
    feed_prof_.clear();
    if (declined) return 1;
    stats_.total_s += now_s() - t_begin;
    return RTJX_OK;
}

}  // namespace rtjx
