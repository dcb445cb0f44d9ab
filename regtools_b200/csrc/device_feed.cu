// regtools_b200/csrc/device_feed.cu — BAM record split + field extraction on the device.
//
// Input: the inflated BGZF payload of a chunk of the file (inflate.cu), i.e. a byte stream of BAM
// records `block_size:i32, core[32], qname, cigar[n_cigar], seq, qual, aux`
// (/root/reference/src/utils/htslib/sam.c:399-432 bam_read1).  Output: the SoA alignment batch the
// cigar_scan kernel consumes (tid, pos, meta, cig_off, cigar) — the same arrays the host feeder
// (bam_feeder.cc) produces, bit for bit.
//
// Record boundaries form a sequential chain (each record's length sits in its first 4 bytes), so the
// stream is cut at SEEDS — virtual offsets taken from the BAI (linear-index entries and bin-chunk
// starts are record starts by construction, hts.c:1288-1350) — and one thread walks each segment.
// Every walk must land exactly on the next seed; a miss raises a flag and the host falls back to
// its own feeder, so exactness never depends on the index being right.
#include "jx_device.cuh"
#include <cub/device/device_scan.cuh>

namespace rtjx {

__device__ __forceinline__ uint32_t ld_u32_any(const uint8_t* p) {             // unaligned 32-bit load
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
    const uint32_t sh = (uint32_t)(a & 3) * 8;
    const uint32_t lo = w[0];
    if (sh == 0) return lo;
    return __funnelshift_r(lo, w[1], sh);
}
__device__ __forceinline__ uint32_t ld_u16_any(const uint8_t* p) { return (uint32_t)p[0] | (uint32_t)p[1] << 8; }

// bam_read1's validity checks (sam.c:399-432); `core` points at the 32-byte core.
__device__ __forceinline__ bool record_ok(const uint8_t* core, int32_t block_len) {
    const int32_t l_data = block_len - 32;
    const int32_t l_qseq = (int32_t)ld_u32_any(core + 16);
    const uint32_t l_qname = core[8];
    if (l_data < 0 || l_qseq < 0 || l_qname < 1) return false;
    const long long aux_off = (long long)l_qname + 4ll * ld_u16_any(core + 12) + ((long long)l_qseq + 1) / 2 + l_qseq;
    return aux_off <= l_data;
}

// One WARP per seed segment.  data = first byte of the chunk's inflated stream (negative offsets reach into
// the carry headroom).  seeds[i] are offsets relative to data; seeds[0] is replaced by -(carry length) when
// `use_carry`.  rec_off gets, per segment, the offsets of its records.
// The chain of block_size fields is serial, but its latency need not be DRAM latency: the stream is pulled
// through a shared-memory ring of four 2 KB windows with 16-byte cp.async — while lane 0 hops from record
// to record inside window w (two aligned LDS + a funnel shift per hop; w and w+1 are resident, a header may
// straddle them), windows w+2 and w+3 are in flight.  Only block_size is looked at here; bam_read1's consistency checks run in parallel in the gather
// kernel.
constexpr int WALK_WARPS = 4;
constexpr int WALK_WIN   = 2048;                 // bytes per window
constexpr int WALK_NWIN  = 4;                    // windows in the ring (power of two)

__global__ void __launch_bounds__(WALK_WARPS * 32)
record_walk_kernel(const uint8_t* __restrict__ data, int64_t data_len, int64_t limit, const int64_t* __restrict__ seeds,
                   const uint32_t* __restrict__ seg_base, uint32_t n_seg, int use_carry, FeedState* __restrict__ state,
                   int32_t* __restrict__ rec_off, uint32_t* __restrict__ seg_cnt) {
    __shared__ __align__(16) uint32_t s_ring[WALK_WARPS][WALK_NWIN * WALK_WIN / 4];
    const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
    const uint32_t i = blockIdx.x * WALK_WARPS + wib;
    if (i >= n_seg) return;
    int64_t p = seeds[i];
    if (i == 0 && use_carry) p = -(int64_t)state->carry_len;
    // Seeds found on the device (block_seeds_kernel) give every BGZF block one segment; blocks without a record start (the tail
    // of a long record, nothing after the stream's last record) have their seed pulled up to the next block's, data_len at the
    // end.  Such a segment is empty and reports nothing; the segment in front of a run of them that reaches data_len is the
    // stream's last one (it owns the carry).
    if (i > 0 && p >= data_len) { if (lane == 0) seg_cnt[i] = 0; return; }
    const bool last = i + 1 == n_seg || seeds[i + 1] >= data_len;
    const int64_t end = last ? data_len : seeds[i + 1];
    const int64_t stop_at = limit < end ? limit : end;        // range end (contig shard) may cut the last segment
    uint32_t n = 0;
    int32_t* my = rec_off + seg_base[i];
    const uint32_t cap = seg_base[i + 1] - seg_base[i];
    int64_t carry_from = -1;
    uint32_t* ring = s_ring[wib];
    const int64_t padded_len = data_len + 32;                 // the inflated buffer is padded past data_len

    // window w covers stream bytes [A + w*WIN, A + (w+1)*WIN) and lives in ring slot (w % NWIN)
    constexpr uint32_t RING_MASK = WALK_NWIN * WALK_WIN / 4 - 1;
    auto load_window = [&](int64_t A, int64_t w) {
        const int64_t wstart = A + w * WALK_WIN;
        uint32_t* dst = ring + (w & (WALK_NWIN - 1)) * (WALK_WIN / 4);
#pragma unroll
        for (int k = 0; k < WALK_WIN / 16 / 32; ++k) {
            const int64_t o = wstart + ((int64_t)k * 32 + lane) * 16;
            if (o + 16 <= padded_len)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst + (k * 32 + lane) * 4)), "l"(data + o) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    bool done = p >= stop_at;
    while (!done) {
        // (re)start the ring at p: windows 0..3 requested, 0 and 1 awaited
        const int64_t A = (p >> 4) << 4;                      // 16-byte aligned (data itself is 256-byte aligned)
        load_window(A, 0); load_window(A, 1); load_window(A, 2); load_window(A, 3);
        asm volatile("cp.async.wait_group 2;" ::: "memory");
        __syncwarp();
        int64_t w = 0;                                         // window lane 0 is walking; w and w+1 are resident
        bool restart = false;
        while (!done && !restart) {
            int64_t new_p = p;
            int flag = 0;                                      // 1 done, 2 window exhausted, 3 far jump (restart the ring)
            if (lane == 0) {
                const int64_t wend = A + (w + 1) * WALK_WIN;   // records STARTING before wend are handled in this window
                while (true) {
                    if (new_p >= stop_at) { flag = 1; break; }
                    if (new_p >= wend) { flag = new_p >= wend + WALK_WIN ? 3 : 2; break; }
                    if (new_p + 4 > data_len) { carry_from = new_p; flag = 1; break; }
                    const uint32_t bi = (uint32_t)(new_p - A);
                    const uint32_t w0 = ring[(bi >> 2) & RING_MASK], w1 = ring[((bi >> 2) + 1) & RING_MASK];
                    const int32_t bl = (int32_t)__funnelshift_r(w0, w1, (bi & 3u) * 8u);
                    if (bl < 32) { atomicMin(&state->bad_offset, (long long)new_p); flag = 1; break; }      // malformed: iteration ends here
                    if (new_p + 4 + (int64_t)bl > data_len) { carry_from = new_p; flag = 1; break; }         // continues in the next chunk
                    if (n < cap) my[n] = (int32_t)new_p;
                    else atomicOr(&state->flags, FEED_FLAG_CAPACITY);
                    ++n;
                    new_p += 4 + (int64_t)bl;
                }
            }
            flag = __shfl_sync(0xffffffffu, flag, 0);
            p = __shfl_sync(0xffffffffu, new_p, 0);
            if (flag == 1) done = true;
            else if (flag == 3) restart = true;
            else {                                             // advance: window w is consumed, its slot takes w+4
                load_window(A, w + WALK_NWIN);
                // walking w+1 needs w+1 and w+2 resident: everything but the two newest requests
                asm volatile("cp.async.wait_group 2;" ::: "memory");
                __syncwarp();
                ++w;
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
    if (lane != 0) return;
    seg_cnt[i] = n < cap ? n : cap;
    if (!last) {
        if (carry_from >= 0 || (p != end && p < limit && state->bad_offset > p)) atomicOr(&state->flags, FEED_FLAG_SEED_MISS);
    } else {
        // bytes of an unfinished record are carried into the next chunk's headroom
        state->next_carry_from = carry_from >= 0 ? carry_from : (p < data_len && p >= limit ? data_len : p);
        if (p >= limit) state->reached_limit = 1;
    }
}

// ------------------------------------------------------------------------------------------------
// Record starts found on the device: one seed per BGZF block
// ------------------------------------------------------------------------------------------------
// The BAI gives a record start every 16 kb of REFERENCE, which on deep loci means segments of hundreds of thousands of
// records walked by one lane (record_walk was 12 % of the GPU time in round 1, at 7.5 % warps active).  The stream itself has
// a much finer natural grid — the BGZF blocks, <= 64 KiB each — if the first record start inside each block can be found.
// A warp scans its block from the front, 32 byte offsets per step, for the first offset where a well-formed record begins
// (the checks bam_read1 makes, sam.c:399-432, plus field ranges, plus the same checks on the record that would follow); real
// records are a few hundred bytes, so the true start is met within a handful of steps.  The guess needs no trust: record_walk
// must arrive at every seed exactly from the seed before it (FEED_FLAG_SEED_MISS otherwise, and the run falls back to the
// index seeds, then to the host feeder), and the first seed of a stream is known, so by induction every seed that the chain
// reaches is a true record start.
__device__ __forceinline__ bool plausible_one(const uint8_t* data, int64_t o, int64_t data_len, int32_t n_ref, int64_t* next) {
    if (o + 36 > data_len) return false;
    const uint8_t* r = data + o;
    const int32_t bs = (int32_t)ld_u32_any(r);
    if (bs < 32 || bs > (1 << 27)) return false;
    const int32_t tid = (int32_t)ld_u32_any(r + 4), pos = (int32_t)ld_u32_any(r + 8);
    if (tid < -1 || tid >= n_ref || pos < -1) return false;
    const int32_t mtid = (int32_t)ld_u32_any(r + 24), mpos = (int32_t)ld_u32_any(r + 28);
    if (mtid < -1 || mtid >= n_ref || mpos < -1) return false;
    if (!record_ok(r + 4, bs)) return false;
    const uint32_t l_qname = r[12], n_cigar = ld_u16_any(r + 16), flag = ld_u16_any(r + 18);
    // a start guessed one to three bytes early reads every field shifted: it survives the range checks above when the true
    // fields are small numbers (refID 0, short template lengths), so the fields that cannot be small by accident are checked
    // too — a non-empty printable read name (SAM 1.4: [!-?A-~]{1,254}), no undefined flag bits, and below a CIGAR whose query
    // length equals l_seq (seen before these checks: ~1e-3 false starts per BGZF block, each one a declined run)
    if (l_qname < 2u || (flag & 0xf000u)) return false;
    // what the fixed fields account for must leave a sane amount of aux data: a start guessed a byte or two early reads a
    // block_size that is the true one shifted left, i.e. megabytes too large (seen at a rate of ~1e-3 per BGZF block)
    const int32_t l_qseq = (int32_t)ld_u32_any(r + 20);
    const long long fixed = 32ll + l_qname + 4ll * n_cigar + ((long long)l_qseq + 1) / 2 + l_qseq;
    if ((long long)bs - fixed > (1ll << 20)) return false;
    const int64_t name_end = o + 36 + (int64_t)l_qname;           // one past the NUL of the read name
    if (name_end <= data_len && data[name_end - 1] != 0) return false;
    // read names are printable ([!-~], SAM spec 1.4): checked on the first characters
    for (uint32_t c = 0; c + 1 < l_qname && c < 6u && o + 36 + c < data_len; ++c) {
        const uint32_t ch = data[o + 36 + c];
        if (ch < 33u || ch > 126u) return false;
    }
    if (n_cigar && n_cigar <= 16u && l_qseq > 0 && !(flag & 4u) && name_end + 4ll * n_cigar <= data_len) {
        uint32_t qlen = 0;
        for (uint32_t c = 0; c < n_cigar; ++c) {
            const uint32_t w = ld_u32_any(data + name_end + 4 * c), op = w & 0xfu;
            if (op > 9u) return false;
            if ((0x193u >> op) & 1u) qlen += w >> 4;              // M, I, S, =, X consume the query
        }
        if (qlen != (uint32_t)l_qseq) return false;
    } else if (n_cigar && name_end + 4 <= data_len && (ld_u32_any(data + name_end) & 0xfu) > 9u) return false;
    *next = o + 4 + (int64_t)bs;
    return true;
}
__device__ __forceinline__ bool plausible_record(const uint8_t* data, int64_t o, int64_t data_len, int32_t n_ref) {
    int64_t nxt = 0, nxt2 = 0, nxt3 = 0;
    if (!plausible_one(data, o, data_len, n_ref, &nxt)) return false;
    if (nxt + 36 > data_len) return true;                         // the next record lies (partly) in the next chunk
    if (!plausible_one(data, nxt, data_len, n_ref, &nxt2)) return false;
    if (nxt2 + 36 > data_len) return true;
    return plausible_one(data, nxt2, data_len, n_ref, &nxt3);
}

constexpr int SEED_WARPS = 8;
// blocks[i] = BgzfBlockDesc of block i (out_off / out_len relative to data).  seeds[0] is the caller's (known start, or the
// carry); seeds[i >= 1] = offset of the first plausible record start in block i, LLONG_MAX if there is none.
__global__ void __launch_bounds__(SEED_WARPS * 32)
block_seeds_kernel(const uint8_t* __restrict__ data, int64_t data_len, const BgzfBlockDesc* __restrict__ blocks, uint32_t n_blocks,
                   int32_t n_ref, int64_t* __restrict__ seeds) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t i = blockIdx.x * SEED_WARPS + (threadIdx.x >> 5);
    if (i == 0 || i >= n_blocks) return;
    const int64_t lo = (int64_t)blocks[i].out_off, hi = lo + (int64_t)blocks[i].out_len;
    int64_t found = 0x7fffffffffffffffll;
    for (int64_t o = lo; o < hi; o += 32) {
        const bool ok = o + lane < hi && plausible_record(data, o + lane, data_len, n_ref);
        const uint32_t m = __ballot_sync(0xffffffffu, ok);
        if (m) { found = o + (__ffs(m) - 1); break; }
    }
    if (lane == 0) seeds[i] = found;
}
// seeds[i] = min(seeds[i .. n)) with data_len as the identity: blocks without a record start take the next block's seed.
__global__ void __launch_bounds__(1024)
seed_suffix_min_kernel(int64_t* __restrict__ seeds, uint32_t n, int64_t data_len) {
    __shared__ int64_t part[1024];
    const uint32_t t = threadIdx.x;
    const uint32_t per = (n + 1023u) / 1024u;
    const uint32_t a = min(t * per, n), b = min(a + per, n);
    int64_t m = data_len;
    for (uint32_t i = b; i-- > a;) { const int64_t v = seeds[i]; if (i > 0 && v < m) m = v; }
    part[t] = m;
    __syncthreads();
    // suffix minimum over the threads' partial results (Hillis-Steele, 10 steps)
    for (uint32_t d = 1; d < 1024u; d <<= 1) {
        const int64_t other = t + d < 1024u ? part[t + d] : data_len;
        __syncthreads();
        if (other < part[t]) part[t] = other;
        __syncthreads();
    }
    int64_t run = t + 1 < 1024u ? part[t + 1] : data_len;        // minimum of everything right of this thread's range
    for (uint32_t i = b; i-- > a;) {
        if (i == 0) break;                                         // seeds[0] is the caller's
        const int64_t v = seeds[i];
        if (v < run) run = v;
        seeds[i] = run;
    }
}
void launch_block_seeds(const uint8_t* data, int64_t data_len, const void* blocks, uint32_t n_blocks, int32_t n_ref, int64_t* seeds,
                        cudaStream_t stream) {
    if (n_blocks <= 1) return;
    block_seeds_kernel<<<(n_blocks + SEED_WARPS - 1) / SEED_WARPS, SEED_WARPS * 32, 0, stream>>>(data, data_len, static_cast<const BgzfBlockDesc*>(blocks),
                                                                                               n_blocks, n_ref, seeds);
    seed_suffix_min_kernel<<<1, 1024, 0, stream>>>(seeds, n_blocks, data_len);
}

// Gathers the per-segment record lists into one dense array; also fetches n_cigar for the scan.
__global__ void __launch_bounds__(256)
record_gather_kernel(const uint8_t* __restrict__ data, const int32_t* __restrict__ rec_off, const uint32_t* __restrict__ seg_base,
                     const uint32_t* __restrict__ seg_cnt, const uint32_t* __restrict__ seg_scan, uint32_t n_seg,
                     uint32_t cap_total, FeedState* __restrict__ state, int32_t* __restrict__ dense,
                     uint32_t* __restrict__ ncig) {
    const uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= cap_total) return;
    uint32_t lo = 0, hi = n_seg;                                   // last segment with seg_base <= s
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (seg_base[mid] <= s) lo = mid; else hi = mid; }
    const uint32_t j = s - seg_base[lo];
    if (j >= seg_cnt[lo]) return;
    const int32_t off = rec_off[s];
    const uint32_t r = seg_scan[lo] + j;
    dense[r] = off;
    ncig[r] = ld_u16_any(data + off + 4 + 12);
    // bam_read1's consistency checks (sam.c:399-432); a failing record ends the reference's iteration -> host path
    if (!record_ok(data + off + 4, (int32_t)ld_u32_any(data + off))) atomicMin(&state->bad_offset, (long long)off);
}

// bam_aux_get + bam_aux2A (sam.c:1254-1266, 1301-1307): value byte of the first `tag` if its type is 'A'.
__device__ __forceinline__ uint32_t strand_tag_byte(const uint8_t* s, const uint8_t* e, uint32_t t0, uint32_t t1) {
    while (s + 3 <= e) {
        const bool match = s[0] == t0 && s[1] == t1;
        const uint32_t type = s[2];
        s += 3;
        if (match) return (type == 'A' && s < e) ? *s : 0u;
        switch (type) {
        case 'A': case 'c': case 'C': s += 1; break;
        case 's': case 'S': s += 2; break;
        case 'i': case 'I': case 'f': s += 4; break;
        case 'd': s += 8; break;
        case 'Z': case 'H': while (s < e && *s) ++s; ++s; break;
        case 'B': {
            if (s + 5 > e) return 0u;
            const uint32_t sub = s[0]; const uint32_t n = ld_u32_any(s + 1); s += 5;
            const uint32_t sz = (sub == 'c' || sub == 'C' || sub == 'A') ? 1u : (sub == 's' || sub == 'S') ? 2u :
                                (sub == 'i' || sub == 'I' || sub == 'f') ? 4u : sub == 'd' ? 8u : 0u;
            if ((unsigned long long)(e - s) < (unsigned long long)sz * n) return 0u;
            s += (size_t)sz * n; break; }
        default: return 0u;
        }
    }
    return 0u;
}

// One thread per record: core fields, CIGAR words, strand tag -> SoA batch.
__global__ void __launch_bounds__(256)
record_extract_kernel(const uint8_t* __restrict__ data, const int32_t* __restrict__ dense, const uint32_t* __restrict__ ncig_scan,
                      const uint32_t* __restrict__ d_n_rec, int32_t n_ref, int xs_mode, uint32_t tag0, uint32_t tag1,
                      int32_t* __restrict__ o_tid, int32_t* __restrict__ o_pos, uint32_t* __restrict__ o_meta,
                      uint32_t* __restrict__ o_off, uint32_t* __restrict__ o_cigar, uint32_t rec_cap, uint32_t cigar_cap,
                      FeedState* __restrict__ state) {
    const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n_rec = *d_n_rec;
    if (r > n_rec) return;
    // this group's alignments go behind the ones already accumulated
    const uint32_t rb = state->acc_rec, ob = state->acc_ops;
    if ((unsigned long long)rb + n_rec + 1ull > rec_cap) { if (r == 0) atomicOr(&state->flags, FEED_FLAG_CAPACITY); return; }
    o_tid += rb; o_pos += rb; o_meta += rb; o_off += rb; o_cigar += ob;
    cigar_cap -= ob < cigar_cap ? ob : cigar_cap;
    if (r == n_rec) { o_off[r] = ob + ncig_scan[r]; return; }      // closing offset (scan is over n_rec + 1 entries)
    const uint8_t* rec = data + dense[r];
    const int32_t bl = (int32_t)ld_u32_any(rec);
    const uint8_t* core = rec + 4;
    int32_t tid = (int32_t)ld_u32_any(core);
    const uint32_t l_qname = core[8], mapq = core[9], n_cigar = ld_u16_any(core + 12), flag = ld_u16_any(core + 14);
    const uint8_t* cig = core + 32 + l_qname;
    const uint32_t o0 = ncig_scan[r];
    if ((unsigned long long)o0 + n_cigar > cigar_cap) { atomicOr(&state->flags, FEED_FLAG_CAPACITY); return; }
    uint32_t strand = 0, nn = 0;
    for (uint32_t k = 0; k < n_cigar; ++k) {
        const uint32_t w = ld_u32_any(cig + 4 * k);
        o_cigar[o0 + k] = w;
        nn += (w & 0xfu) == 3u;
    }
    if (n_cigar > 1) {
        if (tid < 0 || tid >= n_ref) tid = -1;
        if (xs_mode && nn) {
            const int32_t l_qseq = (int32_t)ld_u32_any(core + 16);
            const uint8_t* aux = cig + 4 * (size_t)n_cigar + ((size_t)l_qseq + 1) / 2 + (size_t)l_qseq;
            strand = strand_tag_byte(aux, core + bl, tag0, tag1);
        }
        if (nn) atomicAdd(&state->n_junction_ops, nn);
    }
    o_tid[r] = tid;
    o_pos[r] = (int32_t)ld_u32_any(core + 4);
    o_meta[r] = flag << 16 | mapq << 8 | strand;
    o_off[r] = ob + o0;
}

// Moves the unfinished tail of this chunk in front of the next chunk's data and publishes counts.
__global__ void feed_finish_kernel(const uint8_t* __restrict__ data, int64_t data_len, uint8_t* __restrict__ next_data,
                                   uint32_t headroom, const uint32_t* __restrict__ seg_scan, uint32_t n_seg,
                                   const uint32_t* __restrict__ ncig_scan, FeedState* __restrict__ state) {
    __shared__ uint32_t s_len;
    if (threadIdx.x == 0) {
        long long from = state->next_carry_from;
        if (state->bad_offset <= from) from = data_len;             // stream is dead: nothing to carry
        long long len = data_len - from;
        if (len < 0) len = 0;
        if (len > (long long)headroom || len > data_len) { len = 0; state->flags |= FEED_FLAG_CARRY_TOO_BIG; }
        s_len = (uint32_t)len;
    }
    __syncthreads();
    const uint32_t len = s_len;
    const uint8_t* src = data + (data_len - len);
    for (uint32_t k = threadIdx.x; k < len; k += blockDim.x) next_data[(long long)k - (long long)len] = src[k];
    if (threadIdx.x == 0) {
        state->carry_len = len;
        // (n_rec / n_ops were published by feed_count / feed_ops; a group that raised a flag is never scanned)
        state->acc_rec += state->n_rec; state->acc_ops += state->n_ops; state->acc_jops += state->n_junction_ops;
    }
}
__global__ void feed_carry_in_kernel(const uint8_t* __restrict__ carry_end, uint8_t* __restrict__ data, const FeedState* __restrict__ state) {
    const uint32_t len = state->carry_len;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < len; k += gridDim.x * blockDim.x)
        data[(long long)k - (long long)len] = carry_end[(long long)k - (long long)len];
}
void launch_feed_carry_in(const uint8_t* carry_end, uint8_t* data, const FeedState* state, cudaStream_t stream) {
    feed_carry_in_kernel<<<8, 256, 0, stream>>>(carry_end, data, state);
}
__global__ void feed_acc_reset_kernel(FeedState* state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) { state->acc_rec = state->acc_ops = state->acc_jops = 0; }
}
void launch_feed_acc_reset(FeedState* state, cudaStream_t stream) { feed_acc_reset_kernel<<<1, 32, 0, stream>>>(state); }

__global__ void feed_count_kernel(const uint32_t* __restrict__ seg_scan, uint32_t n_seg, FeedState* __restrict__ state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state->n_rec = seg_scan[n_seg];
}
__global__ void feed_ops_kernel(const uint32_t* __restrict__ ncig_scan, FeedState* __restrict__ state) {
    if (threadIdx.x == 0 && blockIdx.x == 0) state->n_ops = ncig_scan[state->n_rec];
}

__global__ void feed_reset_kernel(FeedState* state, int keep_carry) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        state->bad_offset = 0x7fffffffffffffffll; state->next_carry_from = 0; state->flags = 0;
        state->n_rec = state->n_ops = state->n_junction_ops = 0; state->reached_limit = 0;
        if (!keep_carry) state->carry_len = 0;
    }
}
void launch_feed_reset(FeedState* state, int keep_carry, cudaStream_t stream) { feed_reset_kernel<<<1, 32, 0, stream>>>(state, keep_carry); }

size_t feed_scan_workspace_bytes(uint32_t n) {
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return b + 256;
}

void launch_record_walk(const uint8_t* data, int64_t data_len, int64_t limit, const int64_t* seeds, const uint32_t* seg_base,
                        uint32_t n_seg, int use_carry, FeedState* state, int32_t* rec_off, uint32_t* seg_cnt,
                        cudaStream_t stream) {
    if (!n_seg) return;
    record_walk_kernel<<<(n_seg + WALK_WARPS - 1) / WALK_WARPS, WALK_WARPS * 32, 0, stream>>>(data, data_len, limit, seeds, seg_base, n_seg, use_carry, state,
                                                                 rec_off, seg_cnt);
}

// seg_cnt[0..n_seg) -> seg_scan[0..n_seg] (exclusive, with total), gather, ncig -> ncig_scan[0..n_rec] (exclusive, with total)
void launch_record_gather(const uint8_t* data, const int32_t* rec_off, const uint32_t* seg_base, uint32_t* seg_cnt,
                          uint32_t* seg_scan, uint32_t n_seg, uint32_t cap_total, FeedState* state, int32_t* dense,
                          uint32_t* ncig, uint32_t* ncig_scan, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (!n_seg) return;
    // seg_cnt has one extra zero entry so the exclusive scan yields the total in seg_scan[n_seg]
    cub::DeviceScan::ExclusiveSum(ws, ws_bytes, seg_cnt, seg_scan, (int)(n_seg + 1), stream);
    feed_count_kernel<<<1, 32, 0, stream>>>(seg_scan, n_seg, state);
    cudaMemsetAsync(ncig, 0, ((size_t)cap_total + 1) * sizeof(uint32_t), stream);
    record_gather_kernel<<<(cap_total + 255) / 256, 256, 0, stream>>>(data, rec_off, seg_base, seg_cnt, seg_scan, n_seg, cap_total,
                                                                       state, dense, ncig);
    cub::DeviceScan::ExclusiveSum(ws, ws_bytes, ncig, ncig_scan, (int)(cap_total + 1), stream);
    feed_ops_kernel<<<1, 32, 0, stream>>>(ncig_scan, state);
}

void launch_record_extract(const uint8_t* data, const int32_t* dense, const uint32_t* ncig_scan, uint32_t cap_total,
                           FeedState* state, int32_t n_ref, int xs_mode, uint32_t tag0, uint32_t tag1, int32_t* o_tid,
                           int32_t* o_pos, uint32_t* o_meta, uint32_t* o_off, uint32_t* o_cigar, uint32_t rec_cap, uint32_t cigar_cap,
                           cudaStream_t stream) {
    record_extract_kernel<<<(cap_total + 1 + 255) / 256, 256, 0, stream>>>(data, dense, ncig_scan, &state->n_rec, n_ref, xs_mode, tag0, tag1,
                                                                          o_tid, o_pos, o_meta, o_off, o_cigar, rec_cap, cigar_cap, state);
}

void launch_feed_finish(const uint8_t* data, int64_t data_len, uint8_t* next_data, uint32_t headroom, const uint32_t* seg_scan,
                        uint32_t n_seg, const uint32_t* ncig_scan, FeedState* state, cudaStream_t stream) {
    feed_finish_kernel<<<1, 256, 0, stream>>>(data, data_len, next_data, headroom, seg_scan, n_seg, ncig_scan, state);
}

}  // namespace rtjx
