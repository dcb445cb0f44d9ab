// regtools_b200/csrc/kernels.cu — hand-written sm_100a kernels of the junction-extraction hot path.
//
//   cigar_scan      parse_alignment_into_junctions  (/root/reference/src/junctions/junctions_extractor.cc:377-497)
//                   + set_junction_strand{,_XS,_flag} (:345-359, :283-294, :297-322)
//   junction_merge  junction_qc (:160-170) + add_junction (:174-235)
//   table_compact / finalize_sort   create_junctions_vector + sort_junctions + name ranking
//                   (:538-544, junctions_extractor.h:117-146, :152-157)
//
// The reference walks one alignment at a time through std::string / std::map; here a batch of
// alignments is a SoA slab in HBM (16 B per read + 4 B per CIGAR op) streamed once with coalesced
// 128-bit loads, and the map is a device-wide open-addressed hash updated with atomics after a
// shared-memory pre-aggregation per block.  Integer only; HBM-bound; no tensor cores.
#include "jx_device.cuh"
#include <cub/device/device_merge_sort.cuh>
#include <cstdlib>
#include <cstdio>

namespace rtjx {

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

struct K128 { unsigned long long lo, hi; };

// 128-bit compare-and-swap on global memory (ATOMG.E.CAS.128).
__device__ __forceinline__ K128 cas128(void* addr, K128 cmp, K128 val) {
    K128 old;
    asm volatile(
        "{\n\t.reg .b128 c, v, o;\n\t"
        "mov.b128 c, {%2, %3};\n\t"
        "mov.b128 v, {%4, %5};\n\t"
        "atom.relaxed.gpu.global.cas.b128 o, [%6], c, v;\n\t"
        "mov.b128 {%0, %1}, o;\n\t}"
        : "=l"(old.lo), "=l"(old.hi)
        : "l"(cmp.lo), "l"(cmp.hi), "l"(val.lo), "l"(val.hi), "l"(addr)
        : "memory");
    return old;
}
__device__ __forceinline__ K128 ld128_relaxed(const void* addr) {
    K128 v;
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(v.lo), "=l"(v.hi) : "l"(addr) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t mix_key(unsigned long long lo, unsigned long long hi) {
    unsigned long long h = lo * 0x9E3779B97F4A7C15ull;
    h ^= hi * 0xC2B2AE3D27D4EB4Full;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 29;
    return (uint32_t)h;
}

// strand char of a read (junctions_extractor.cc:283-294 XS, :297-322 flag)
__device__ __forceinline__ uint32_t read_strand(uint32_t meta, int strandness) {
    if (strandness == 0) {
        uint32_t b = meta & 0xffu;
        return b ? b : (uint32_t)'?';
    }
    uint32_t flag = meta >> 16;
    uint32_t rev = (flag >> 4) & 1u, mrev = (flag >> 5) & 1u, r1 = (flag >> 6) & 1u, r2 = (flag >> 7) & 1u;
    uint32_t nb = (strandness == 1) ? 1u : 0u;          // !(strandness_-1)
    uint32_t fs = nb ^ r1 ^ rev, ss = nb ^ r2 ^ mrev;
    return fs != ss ? (uint32_t)'?' : (fs ? (uint32_t)'+' : (uint32_t)'-');
}

// Strand from the intron motif (get_splice_site :564-584 + set_junction_strand_intron_motif :325-342 with fai_fetch's
// clipping, faidx.c:386-397).  `prev` is the strand the reference's reused Junction object carries from the previous
// junction of the same alignment (0 before the first): when it is '-', the two 2-mers are reverse-complemented and
// swapped before the comparison.  Returns '+', '-' or '?' ('?' = fall back to the XS tag / flag strand, :352-358).
__device__ __forceinline__ uint32_t comp_base(uint32_t c) {            // common.h:59-83
    return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
}
__device__ __noinline__ uint32_t motif_strand(const ScanParams& p, int32_t tid, uint32_t start, uint32_t end, uint32_t prev,
                                              uint32_t* counters) {
    const unsigned long long len = (uint32_t)tid < p.g_n ? p.g_len[tid] : ~0ull;
    if (len == ~0ull) {                                                // fai_fetch returns NULL -> runtime_error (:553-555)
        if (atomicCAS(&counters[CTR_GENOME_MISS], 0u, (uint32_t)tid + 1u) == 0u) counters[CTR_GENOME_MISS_POS] = start;
        return '?';
    }
    const uint8_t* g = p.genome + p.g_off[tid];
    // "chrom:start+1-start+2" and "chrom:end-1-end", 1-based inclusive -> [b, e) clipped to the sequence
    unsigned long long b1 = start, e1 = (unsigned long long)start + 2ull;
    unsigned long long b2 = end >= 2u ? end - 2u : 0u, e2 = end;
    b1 = b1 < len ? b1 : len; e1 = e1 < len ? e1 : len; b2 = b2 < len ? b2 : len; e2 = e2 < len ? e2 : len;
    if (e1 - b1 != 2ull || e2 < b2 || e2 - b2 != 2ull) return '?';   // a shorter string equals no 5-character motif
    uint32_t m0 = g[b1], m1 = g[b1 + 1], m2 = g[b2], m3 = g[b2 + 1];
    if (prev == '-') {
        const uint32_t a0 = comp_base(m3), a1 = comp_base(m2), a2 = comp_base(m1), a3 = comp_base(m0);
        m0 = a0; m1 = a1; m2 = a2; m3 = a3;
    }
    const uint32_t code = m0 << 24 | m1 << 16 | m2 << 8 | m3;
    if (code == 0x47544147u /*GTAG*/ || code == 0x47434147u /*GCAG*/ || code == 0x41544143u /*ATAC*/) return '+';
    if (code == 0x43544143u /*CTAC*/ || code == 0x43544743u /*CTGC*/ || code == 0x47544154u /*GTAT*/) return '-';
    return '?';
}

// ------------------------------------------------------------------------------------------------
// cigar_scan
// ------------------------------------------------------------------------------------------------
// One block = one tile of SCAN_TILE consecutive alignments.  Thread t owns alignments
// base + t + j*SCAN_THREADS (j < SCAN_RPT): every metadata load of a warp is one fully coalesced
// 128-byte request, and all 4*SCAN_RPT loads of a thread are issued before the first use.  The
// tile's CIGAR ops are one contiguous slab of the `cigar` array; it is staged into shared memory
// with 128-bit streaming loads and walked from there.  Candidates are staged in shared memory and
// flushed with one global atomicAdd per block and coalesced 128-bit stores.
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_RPT     = 4;
constexpr int SCAN_TILE    = SCAN_THREADS * SCAN_RPT;   // 1024 alignments
constexpr int SCAN_SLAB    = 6144;                      // CIGAR words staged per tile (24 KB)
constexpr int SCAN_STAGE   = 512;                       // candidates staged per tile (16 KB)

struct ScanSmem {
    uint32_t slab[SCAN_SLAB];
    uint4    stage[SCAN_STAGE * 2];
    uint32_t n_stage;
    uint32_t flush_base;
};

__device__ __forceinline__ void scan_emit(ScanSmem& sm, Cand* __restrict__ out, uint32_t cap, uint32_t* counters,
                                          uint32_t start, uint32_t end, uint32_t left, uint32_t right,
                                          uint64_t ord, int32_t tid, uint32_t strand) {
    uint4 a = make_uint4(start, end, start - left, end + right);
    uint4 b = make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)tid, strand);
    uint32_t i = atomicAdd(&sm.n_stage, 1u);
    if (i < SCAN_STAGE) {
        sm.stage[2 * i] = a;
        sm.stage[2 * i + 1] = b;
    } else {                                   // tile denser than the staging buffer: go to HBM directly
        uint32_t g = atomicAdd(&counters[CTR_NCAND], 1u);
        if (g < cap) {
            uint4* o = reinterpret_cast<uint4*>(out + g);
            o[0] = a; o[1] = b;
        } else {
            atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
        }
    }
}

// Closed form of the reference's per-op state machine: for the N op k,
//   start = pos + sum(len of M,=,D,X,N before k),  end = start + len_k,
//   left  = sum(len of M,=) since the last of {N,D,X,I,S},  right likewise up to the next one.
// H, P, B and op codes 10..15 change nothing.
template <bool FROM_SMEM>
__device__ __forceinline__ void scan_walk(ScanSmem& sm, const uint32_t* __restrict__ ops, uint32_t n,
                                          uint32_t pos, int32_t tid, uint32_t strand, uint64_t read_ord,
                                          Cand* __restrict__ out, uint32_t cap, uint32_t* counters) {
    uint32_t cur = pos, run = 0;
    bool pending = false;
    uint32_t p_start = 0, p_end = 0, p_left = 0, p_k = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t w = FROM_SMEM ? ops[i] : __ldg(ops + i);
        uint32_t op = w & 0xfu, len = w >> 4;
        // bit masks over op codes: M=0 I=1 D=2 N=3 S=4 H=5 P=6 '='=7 X=8 B=9
        const uint32_t ANC = (1u << 0) | (1u << 7);
        const uint32_t BRK = (1u << 3) | (1u << 2) | (1u << 8) | (1u << 1) | (1u << 4);
        const uint32_t REFC = ANC | (1u << 2) | (1u << 8) | (1u << 3);
        uint32_t bit = 1u << op;
        if (bit & BRK) {
            if (pending) {
                scan_emit(sm, out, cap, counters, p_start, p_end, p_left, run,
                          read_ord << 16 | (p_k > 0xffffu ? 0xffffu : p_k), tid, strand);
                pending = false;
            }
            if (op == 3u) {
                pending = true; p_start = cur; p_end = cur + len; p_left = run; p_k = i;
            }
            run = 0;
        } else if (bit & ANC) {
            run += len;
        }
        if (bit & REFC) cur += len;
    }
    if (pending)
        scan_emit(sm, out, cap, counters, p_start, p_end, p_left, run,
                  read_ord << 16 | (p_k > 0xffffu ? 0xffffu : p_k), tid, strand);
}

__global__ void __launch_bounds__(SCAN_THREADS, 4)
cigar_scan_kernel(BatchView b, ScanParams prm, Cand* __restrict__ out, uint32_t cap, uint32_t* __restrict__ counters) {
    __shared__ ScanSmem sm;
    const uint32_t t = threadIdx.x;
    const uint32_t base = blockIdx.x * SCAN_TILE;
    const uint32_t n_tile = min((uint32_t)SCAN_TILE, b.n_reads - base);

    // ---- phase 1: issue every metadata load of this thread (coalesced, streaming) ----
    uint32_t off0[SCAN_RPT], off1[SCAN_RPT], pos[SCAN_RPT], meta[SCAN_RPT];
    int32_t tid[SCAN_RPT];
#pragma unroll
    for (int j = 0; j < SCAN_RPT; ++j) {
        uint32_t r = t + j * SCAN_THREADS;
        bool ok = r < n_tile;
        uint32_t i = base + (ok ? r : 0);
        off0[j] = __ldg(b.cig_off + i);
        off1[j] = __ldg(b.cig_off + i + 1);          // same lines as off0: merged in L1
        pos[j]  = ldg_stream_u32(reinterpret_cast<const uint32_t*>(b.pos) + i);
        meta[j] = ldg_stream_u32(b.meta + i);
        tid[j]  = (int32_t)ldg_stream_u32(reinterpret_cast<const uint32_t*>(b.tid) + i);
        if (!ok) { off1[j] = off0[j]; }
    }
    if (t == 0) sm.n_stage = 0;

    // ---- phase 2: stage the tile's CIGAR slab ----
    const uint32_t slab_lo = __ldg(b.cig_off + base);
    const uint32_t slab_hi = __ldg(b.cig_off + base + n_tile);
    const uint32_t a0 = slab_lo & ~3u;                        // 16-byte aligned start (word index)
    const bool staged = (slab_hi - a0) <= (uint32_t)SCAN_SLAB;
    if (staged) {
        const uint32_t n_vec = (slab_hi - a0 + 3u) >> 2;
        const uint32_t full_vec = b.n_ops >> 2;               // vectors that lie fully inside the array
        const uint4* src = reinterpret_cast<const uint4*>(b.cigar) + (a0 >> 2);
        uint4* dst = reinterpret_cast<uint4*>(sm.slab);
        for (uint32_t v = t; v < n_vec; v += SCAN_THREADS) {
            if ((a0 >> 2) + v < full_vec) {
                dst[v] = ldg_stream_u4(src + v);
            } else {                                          // ragged tail of the array
                uint32_t w0 = a0 + 4 * v;
                uint4 x;
                x.x = w0 + 0 < b.n_ops ? __ldg(b.cigar + w0 + 0) : 0u;
                x.y = w0 + 1 < b.n_ops ? __ldg(b.cigar + w0 + 1) : 0u;
                x.z = w0 + 2 < b.n_ops ? __ldg(b.cigar + w0 + 2) : 0u;
                x.w = w0 + 3 < b.n_ops ? __ldg(b.cigar + w0 + 3) : 0u;
                dst[v] = x;
            }
        }
    }
    __syncthreads();

    // ---- phase 3: walk the (few) multi-op alignments ----
#pragma unroll
    for (int j = 0; j < SCAN_RPT; ++j) {
        uint32_t n = off1[j] - off0[j];
        if (n > 1u && tid[j] >= 0) {                          // junctions_extractor.cc:379
            uint64_t read_ord = b.first_ordinal + base + t + j * SCAN_THREADS;
            uint32_t strand = read_strand(meta[j], prm.strandness);
            if (staged)
                scan_walk<true>(sm, sm.slab + (off0[j] - a0), n, pos[j], tid[j], strand, read_ord, out, cap, counters);
            else
                scan_walk<false>(sm, b.cigar + off0[j], n, pos[j], tid[j], strand, read_ord, out, cap, counters);
        }
    }
    __syncthreads();

    // ---- phase 4: flush staged candidates, one reservation per block ----
    const uint32_t n_st = min(sm.n_stage, (uint32_t)SCAN_STAGE);
    if (n_st == 0) return;
    if (t == 0) sm.flush_base = atomicAdd(&counters[CTR_NCAND], n_st);
    __syncthreads();
    const uint32_t fb = sm.flush_base;
    uint4* o = reinterpret_cast<uint4*>(out);
    for (uint32_t v = t; v < 2 * n_st; v += SCAN_THREADS) {
        uint32_t c = fb + (v >> 1);
        if (c < cap) o[2ull * fb + v] = sm.stage[v];
        else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
    }
}

static int num_sms();

// ---- mbarrier / bulk-copy primitives (used by the warp-specialised variant) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// 1-D bulk async copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}


// ------------------------------------------------------------------------------------------------
// cigar_scan, warp-specialised persistent version (variant 4, the one normally launched)
// ------------------------------------------------------------------------------------------------
// 8 consumer warps + 1 service warp per CTA, two CTAs per SM, tiles of 1024 alignments strided
// over the grid.  The service warp is the only one that talks to HBM: it keeps a 3-stage ring of
// tile metadata (pos, meta, tid, cig_off: 16 KB) and CIGAR slabs full with 1-D bulk async copies
// (cp.async.bulk -> UBLKCP) that complete on mbarriers, and it drains the double-buffered
// candidate staging area with one global reservation per tile.  Consumers never wait for a load
// they could have been told about earlier, never wait for a flush, and synchronise among
// themselves once per tile (named barrier after the compaction pass).
constexpr int S4_CONSUMERS = 256;
constexpr int S4_THREADS   = S4_CONSUMERS + 32;
constexpr int S4_TILE      = 1024;
constexpr int S4_STAGES    = 3;
constexpr int S4_SLAB      = 2560;               // CIGAR words per stage (10 KB)
constexpr int S4_OUT       = 384;                // staged candidates per buffer (12 KB), two buffers

struct alignas(16) S4Stage {
    uint32_t pos[S4_TILE];
    uint32_t meta[S4_TILE];
    uint32_t tid[S4_TILE];
    uint32_t off[S4_TILE + 4];
    uint32_t slab[S4_SLAB];
};
struct alignas(16) S4Smem {
    S4Stage st[S4_STAGES];
    uint4 out[2][S4_OUT * 2];
    unsigned long long meta_full[S4_STAGES];
    unsigned long long slab_full[S4_STAGES];
    unsigned long long tile_done[2];
    unsigned long long out_free[2];
    uint32_t slab_a0[S4_STAGES];
    uint32_t slab_direct[S4_STAGES];
    uint32_t n_work[2], n_out[2];
    uint16_t work[2][S4_TILE];
};

__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
constexpr int S4_L2_AHEAD = 3;                   // tiles prefetched into L2 beyond the shared-memory ring

struct S4Emit {
    S4Smem& sm; uint32_t buf; Cand* __restrict__ out; uint32_t cap; uint32_t* counters; uint32_t dbg;
    __device__ __forceinline__ void operator()(uint32_t start, uint32_t end, uint32_t left, uint32_t right,
                                               uint64_t ord, int32_t tid, uint32_t strand) const {
        uint4 a = make_uint4(start, end, start - left, end + right);
        uint4 b = make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)tid, strand);
        if (dbg & 4u) { if ((a.x ^ a.y ^ a.z ^ a.w ^ b.x) == 0x9e3779b9u) sm.n_out[buf] = 1; return; }
        // one shared-memory atomic per converged group of lanes instead of one per lane
        const uint32_t mask = __activemask();
        const uint32_t lane = threadIdx.x & 31u;
        const int leader = __ffs(mask) - 1;
        uint32_t i = 0;
        if ((int)lane == leader) i = atomicAdd(&sm.n_out[buf], (uint32_t)__popc(mask));
        i = __shfl_sync(mask, i, leader) + __popc(mask & ((1u << lane) - 1u));
        if (i < S4_OUT) {
            sm.out[buf][2 * i] = a; sm.out[buf][2 * i + 1] = b;
        } else {
            uint32_t g = atomicAdd(&counters[CTR_NCAND], 1u);
            if (g < cap) { uint4* o = reinterpret_cast<uint4*>(out + g); o[0] = a; o[1] = b; }
            else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
        }
    }
};

// Walk with the first 8 ops preloaded by independent loads and a predicated (branch-free) state
// update; only the emits diverge.  Same arithmetic as walk_lin; op code 15 is a transparent filler.
template <bool FROM_SMEM, class Emit>
__device__ __forceinline__ void walk_fast(const uint32_t* __restrict__ ops, uint32_t n, uint32_t pos, int32_t tid,
                                          uint32_t strand, uint64_t read_ord, const Emit& emit) {
    const uint32_t ANC = (1u << 0) | (1u << 7);
    const uint32_t BRK = (1u << 3) | (1u << 2) | (1u << 8) | (1u << 1) | (1u << 4);
    const uint32_t REFC = ANC | (1u << 2) | (1u << 8) | (1u << 3);
    constexpr int PRE = 4;                        // ops preloaded by independent loads (covers 50M100N50M, 5S45M100N50M, ...)
    uint32_t w[PRE];
#pragma unroll
    for (int i = 0; i < PRE; ++i) w[i] = (uint32_t)i < n ? (FROM_SMEM ? ops[i] : __ldg(ops + i)) : 0xfu;
    uint32_t cur = pos, run = 0;
    bool pending = false;
    uint32_t p_start = 0, p_end = 0, p_left = 0, p_k = 0;
#pragma unroll
    for (int i = 0; i < PRE; ++i) {
        const uint32_t op = w[i] & 0xfu, len = w[i] >> 4, bit = 1u << op;
        const bool brk = (bit & BRK) != 0, is_n = op == 3u;
        if (pending && brk) emit(p_start, p_end, p_left, run, read_ord << 16 | p_k, tid, strand);
        if (brk) pending = is_n;
        if (is_n) { p_start = cur; p_end = cur + len; p_left = run; p_k = (uint32_t)i; }
        run = brk ? 0u : run + ((bit & ANC) ? len : 0u);
        cur += (bit & REFC) ? len : 0u;
    }
    for (uint32_t i = PRE; i < n; ++i) {
        const uint32_t x = FROM_SMEM ? ops[i] : __ldg(ops + i);
        const uint32_t op = x & 0xfu, len = x >> 4, bit = 1u << op;
        const bool brk = (bit & BRK) != 0, is_n = op == 3u;
        if (pending && brk) emit(p_start, p_end, p_left, run, read_ord << 16 | (p_k > 0xffffu ? 0xffffu : p_k), tid, strand);
        if (brk) pending = is_n;
        if (is_n) { p_start = cur; p_end = cur + len; p_left = run; p_k = i; }
        run = brk ? 0u : run + ((bit & ANC) ? len : 0u);
        cur += (bit & REFC) ? len : 0u;
    }
    if (pending) emit(p_start, p_end, p_left, run, read_ord << 16 | (p_k > 0xffffu ? 0xffffu : p_k), tid, strand);
}

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, %0;" ::"n"(S4_CONSUMERS) : "memory"); }

// one warp copies the staged candidates of buffer `buf` to HBM with a single reservation
__device__ __forceinline__ void s4_flush(S4Smem& sm, uint32_t buf, uint32_t lane, Cand* __restrict__ out, uint32_t cap,
                                         uint32_t* counters) {
    const uint32_t n_st = min(sm.n_out[buf], (uint32_t)S4_OUT);
    __syncwarp();
    if (n_st) {
        uint32_t fb = 0;
        if (lane == 0) fb = atomicAdd(&counters[CTR_NCAND], n_st);
        fb = __shfl_sync(0xffffffffu, fb, 0);
        uint4* o = reinterpret_cast<uint4*>(out);
        for (uint32_t v = lane; v < 2 * n_st; v += 32) {
            if (fb + (v >> 1) < cap) o[2ull * fb + v] = sm.out[buf][v];
            else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
        }
    }
    __syncwarp();
    if (lane == 0) sm.n_out[buf] = 0;
}

__global__ void __launch_bounds__(S4_THREADS, 2)
cigar_scan_ws_kernel(BatchView b, ScanParams prm, Cand* __restrict__ out, uint32_t cap, uint32_t* __restrict__ counters) {
    extern __shared__ __align__(128) unsigned char s4_raw[];
    S4Smem& sm = *reinterpret_cast<S4Smem*>(s4_raw);
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t n_tiles = (b.n_reads + S4_TILE - 1) / S4_TILE;
    const uint32_t n_ops_vec_end = b.n_ops & ~3u;
    const uint32_t n_my = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    auto tile_of = [&](uint32_t kk) { return blockIdx.x + kk * gridDim.x; };

    if (t == 0) {
        for (int s = 0; s < S4_STAGES; ++s) { mbar_init(&sm.meta_full[s], 1); mbar_init(&sm.slab_full[s], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&sm.tile_done[i], S4_CONSUMERS / 32); mbar_init(&sm.out_free[i], 1); }
        sm.n_work[0] = sm.n_work[1] = 0; sm.n_out[0] = sm.n_out[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (t >= (uint32_t)S4_CONSUMERS) {
        // ======================= service warp: loads and stores =======================
        auto issue_meta = [&](uint32_t tile, int s) {
            const uint32_t base = tile * S4_TILE;
            S4Stage& st = sm.st[s];
            if (base + S4_TILE + 3 <= b.n_reads) {
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    mbar_arrive_expect_tx(&sm.meta_full[s], 3 * S4_TILE * 4 + (S4_TILE + 4) * 4);
                    bulk_g2s(st.off, b.cig_off + base, (S4_TILE + 4) * 4, &sm.meta_full[s]);
                    bulk_g2s(st.pos, b.pos + base, S4_TILE * 4, &sm.meta_full[s]);
                    bulk_g2s(st.meta, b.meta + base, S4_TILE * 4, &sm.meta_full[s]);
                    bulk_g2s(st.tid, b.tid + base, S4_TILE * 4, &sm.meta_full[s]);
                }
            } else {                                           // ragged tail of the batch: plain loads
                const uint32_t n_tile = min((uint32_t)S4_TILE, b.n_reads - base);
                for (uint32_t r = lane; r < n_tile; r += 32) {
                    st.pos[r] = (uint32_t)b.pos[base + r]; st.meta[r] = b.meta[base + r]; st.tid[r] = (uint32_t)b.tid[base + r];
                }
                for (uint32_t r = lane; r <= n_tile; r += 32) st.off[r] = b.cig_off[base + r];
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.meta_full[s]);
            }
        };
        auto issue_slab = [&](uint32_t kk) {                  // lane 0; tile index kk of this CTA
            const int s = (int)(kk % S4_STAGES);
            mbar_wait(&sm.meta_full[s], (kk / S4_STAGES) & 1u);
            S4Stage& st = sm.st[s];
            const uint32_t n_tile = min((uint32_t)S4_TILE, b.n_reads - tile_of(kk) * S4_TILE);
            const uint32_t lo = st.off[0], hi = st.off[n_tile];
            const uint32_t a0 = lo & ~3u, end4 = (hi + 3u) & ~3u;
            sm.slab_a0[s] = a0;
            if (hi <= lo) {
                sm.slab_direct[s] = 0; mbar_arrive(&sm.slab_full[s]);
            } else if (end4 - a0 > (uint32_t)S4_SLAB || end4 > n_ops_vec_end) {
                sm.slab_direct[s] = 1; mbar_arrive(&sm.slab_full[s]);
            } else {
                sm.slab_direct[s] = 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(&sm.slab_full[s], (end4 - a0) * 4);
                bulk_g2s(st.slab, b.cigar + a0, (end4 - a0) * 4, &sm.slab_full[s]);
            }
        };
        // L2 prefetch of a tile's metadata columns (lanes 0..3 take one column each): keeps more bytes in
        // flight to HBM than the shared-memory ring alone can hold
        auto prefetch_meta = [&](uint32_t kk) {
            if (kk >= n_my || (prm.debug & 16u)) return;
            const uint32_t base = tile_of(kk) * S4_TILE;
            if (base + S4_TILE + 3 > b.n_reads) return;
            if (lane == 0) bulk_prefetch_l2(b.cig_off + base, (S4_TILE + 4) * 4);
            else if (lane == 1) bulk_prefetch_l2(b.pos + base, S4_TILE * 4);
            else if (lane == 2) bulk_prefetch_l2(b.meta + base, S4_TILE * 4);
            else if (lane == 3) bulk_prefetch_l2(b.tid + base, S4_TILE * 4);
        };
        for (uint32_t i = 0; i < (uint32_t)S4_STAGES && i < n_my; ++i) issue_meta(tile_of(i), (int)i);
        for (uint32_t i = 0; i < (uint32_t)S4_L2_AHEAD; ++i) prefetch_meta(S4_STAGES + i);
        if (lane == 0) for (uint32_t i = 0; i < (uint32_t)S4_STAGES && i < n_my; ++i) issue_slab(i);
        for (uint32_t k = 0; k < n_my; ++k) {
            const uint32_t cb = k & 1u;
            mbar_wait(&sm.tile_done[cb], (k / 2) & 1u);        // every consumer warp is past tile k
            if (k + S4_STAGES < n_my) issue_meta(tile_of(k + S4_STAGES), (int)(k % S4_STAGES));
            prefetch_meta(k + S4_STAGES + S4_L2_AHEAD);
            if (prm.debug & 8u) { if (lane == 0) sm.n_out[cb] = 0; } else s4_flush(sm, cb, lane, out, cap, counters);
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&sm.out_free[cb]);
                // the slab of tile k+3 goes out as soon as its cig_off column (issued above) has landed:
                // the dependent load gets ~two tiles of slack before consumers need it
                if (k + S4_STAGES < n_my) issue_slab(k + S4_STAGES);
            }
            __syncwarp();
        }
        return;
    }

    // ============================ consumer warps ============================
    for (uint32_t k = 0; k < n_my; ++k) {
        const uint32_t tile = tile_of(k);
        const int s = (int)(k % S4_STAGES);
        const uint32_t parity = (k / S4_STAGES) & 1u;
        const uint32_t cb = k & 1u;
        mbar_wait(&sm.meta_full[s], parity);
        S4Stage& st = sm.st[s];
        const uint32_t base = tile * S4_TILE;
        const uint32_t n_tile = min((uint32_t)S4_TILE, b.n_reads - base);

        // ---- phase A: compact the alignments that have more than one CIGAR op (junctions_extractor.cc:379)
        {
            const uint4 o = *reinterpret_cast<const uint4*>(&st.off[4 * t]);
            const uint32_t o4 = st.off[4 * t + 4];
            const uint32_t r0 = 4 * t;
            uint32_t flags = 0;
            if (r0 + 0 < n_tile && o.y - o.x > 1u) flags |= 1u;
            if (r0 + 1 < n_tile && o.z - o.y > 1u) flags |= 2u;
            if (r0 + 2 < n_tile && o.w - o.z > 1u) flags |= 4u;
            if (r0 + 3 < n_tile && o4 - o.w > 1u) flags |= 8u;
            const uint32_t cnt = __popc(flags);
            uint32_t x = cnt;
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, dlt); if ((int)lane >= dlt) x += y; }
            uint32_t wbase = 0;
            if (lane == 31 && x) wbase = atomicAdd(&sm.n_work[cb], x);
            wbase = __shfl_sync(0xffffffffu, wbase, 31);
            uint32_t p = wbase + x - cnt;
            uint16_t* work = sm.work[cb];
            if (flags & 1u) work[p++] = (uint16_t)(r0 + 0);
            if (flags & 2u) work[p++] = (uint16_t)(r0 + 1);
            if (flags & 4u) work[p++] = (uint16_t)(r0 + 2);
            if (flags & 8u) work[p++] = (uint16_t)(r0 + 3);
        }
        consumer_bar();
        if (t == 0) sm.n_work[cb ^ 1u] = 0;                   // for tile k+1; nobody reads it any more

        // ---- phase B: walk the compacted alignments, one per thread, in rounds of 256
        {
            const uint32_t n_work = (prm.debug & 1u) ? 0u : sm.n_work[cb];
            // unconditional: these waits also keep consumers at most two tiles ahead of the service warp,
            // which the 1-bit phase parity of tile_done / slab_full relies on
            mbar_wait(&sm.slab_full[s], parity);
            if (k >= 2) mbar_wait(&sm.out_free[cb], ((k / 2) - 1u) & 1u);       // staging buffer drained (tile k-2)
            const uint32_t a0 = sm.slab_a0[s];
            const bool direct = sm.slab_direct[s] != 0;
            const S4Emit emit{sm, cb, out, cap, counters, prm.debug};
            const uint16_t* work = sm.work[cb];
            for (uint32_t w0 = 0; w0 < n_work; w0 += S4_CONSUMERS) {
                const uint32_t w = w0 + t;
                if (w < n_work) {
                    const uint32_t r = work[w];
                    const int32_t tid = (int32_t)st.tid[r];
                    if (tid >= 0) {
                        const uint32_t o0 = st.off[r], n = st.off[r + 1] - o0;
                        const uint32_t strand = read_strand(st.meta[r], prm.strandness);
                        const uint64_t read_ord = b.first_ordinal + base + r;
                        if (!direct) walk_fast<true>(st.slab + (o0 - a0), n, st.pos[r], tid, strand, read_ord, emit);
                        else walk_fast<false>(b.cigar + o0, n, st.pos[r], tid, strand, read_ord, emit);
                    }
                }
                if (w0 + S4_CONSUMERS < n_work) {              // dense tile: drain the staging buffer between rounds
                    consumer_bar();
                    if (t < 32) s4_flush(sm, cb, lane, out, cap, counters);
                    consumer_bar();
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.tile_done[cb]);
    }
}

// ------------------------------------------------------------------------------------------------
// cigar_scan, many-small-blocks version (variant 5)
// ------------------------------------------------------------------------------------------------
// 128 threads per block, one tile of 512 alignments per block, ~18 KB of shared memory: up to 12
// blocks are resident per SM and the hardware block scheduler overlaps their phases, so while some
// blocks wait for their (data-dependent) CIGAR slab others are streaming metadata or walking.
// Columns and slab go global -> shared with 16-byte cp.async (LDGSTS), no register staging.
// Template parameters: threads per block (tile = 4 alignments per thread), CIGAR words staged per tile,
// candidates staged per tile.  The default instantiation is <128, 1024, 192>: 512-alignment tiles, 19.5 KB.
template <int S5_THREADS, int S5_SLAB, int S5_OUT>
struct alignas(16) S5SmemT {
    static constexpr int S5_TILE = S5_THREADS * 4;
    uint32_t off[S5_TILE + 4];
    uint32_t pos[S5_TILE];
    uint32_t meta[S5_TILE];
    uint32_t tid[S5_TILE];
    uint32_t slab[S5_SLAB];
    uint4 out[S5_OUT * 2];
    uint32_t n_work, n_out;
    uint16_t work[S5_TILE];
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <class Smem, int S5_OUT, bool MOTIF, bool VREG, bool BC = false>
struct S5Emit {
    Smem& sm; Cand* __restrict__ out; uint32_t cap; uint32_t* counters; uint32_t dbg;
    const ScanParams* prm; uint32_t* jstrand;                 // intron-motif mode: parameters + the alignment's running strand
    const int32_t* rspan;                                     // variant-region mode: [pos, endpos) of the alignment being walked
    __device__ __forceinline__ void push(const uint4& a, const uint4& b) const {
        // one shared-memory atomic per candidate: emit sites are divergent (and the variant-region loop has per-lane trip
        // counts), where __activemask() does not promise that the lanes it names are converged at a following shuffle
        const uint32_t i = atomicAdd(&sm.n_out, 1u);
        if (i < S5_OUT) {
            sm.out[2 * i] = a; sm.out[2 * i + 1] = b;
        } else {
            uint32_t g = atomicAdd(&counters[CTR_NCAND], 1u);
            if (g < cap) { uint4* o = reinterpret_cast<uint4*>(out + g); o[0] = a; o[1] = b; }
            else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
        }
    }
    __device__ __forceinline__ void operator()(uint32_t start, uint32_t end, uint32_t left, uint32_t right,
                                               uint64_t ord, int32_t tid, uint32_t strand) const {
        if (MOTIF) {                                           // set_junction_strand with a FASTA (:345-359): motif first
            const uint32_t m = motif_strand(*prm, tid, start, end, *jstrand, counters);
            if (BC) {                                          // bits 8.. of `strand` carry the barcode id
                if (m != '?') strand = (strand & ~0xffu) | m;
                *jstrand = strand & 0xffu;
            } else {
                if (m != '?') strand = m;
                *jstrand = strand;
            }
        }
        const uint4 a = make_uint4(start, end, start - left, end + right);
        const uint4 b = make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)tid, strand);
        if (dbg & 4u) { if ((a.x ^ a.y ^ a.z ^ a.w ^ b.x) == 0x9e3779b9u) sm.n_out = 1; return; }
        if (!VREG) { push(a, b); return; }
        // one candidate per variant region the ALIGNMENT belongs to (tid, pos < end, endpos > beg; hts.c:1941-1963)
        const VariantRegions& vr = prm->vr;
        const int32_t rp = rspan[0], re = rspan[1];
        uint32_t lo = 0, hi = vr.n;                            // first region with (tid, beg) >= (tid, endpos)
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const int32_t mt = vr.tid[mid];
            if (mt < tid || (mt == tid && vr.beg[mid] < re)) lo = mid + 1; else hi = mid;
        }
        for (uint32_t i = lo; i-- > 0;) {
            if (vr.tid[i] != tid) break;
            if ((long long)vr.beg[i] + (long long)vr.max_len <= (long long)rp) break;    // no earlier region can reach pos
            if (vr.end[i] > rp) push(a, make_uint4(b.x, b.y, b.z, strand | (i + 1u) << 8));
        }
    }
};

// MINB > 1 asks ptxas for that many resident blocks per SM (A/B configurations 8-10, see launch_cigar_scan)
template <int S5_THREADS, int S5_SLAB, int S5_OUT, bool MOTIF = false, bool VREG = false, bool BC = false, int MINB = 1>
__global__ void __launch_bounds__(S5_THREADS, MINB)
cigar_scan_small_kernel(BatchView b, ScanParams prm, Cand* __restrict__ out, uint32_t cap, uint32_t* __restrict__ counters,
                        const uint32_t* __restrict__ tile_off, CandRegions rg) {
    using Smem = S5SmemT<S5_THREADS, S5_SLAB, S5_OUT>;
    constexpr int S5_TILE = Smem::S5_TILE;
    __shared__ Smem sm;
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t base = blockIdx.x * S5_TILE;
    const uint32_t n_tile = min((uint32_t)S5_TILE, b.n_reads - base);
    const bool full = base + S5_TILE + 3 <= b.n_reads;        // the 516-entry cig_off window is in bounds

    // ---- metadata columns: global -> shared, 16 bytes per cp.async
    if (full) {
        cp_async16(&sm.off[4 * t], b.cig_off + base + 4 * t);
        cp_async16(&sm.pos[4 * t], b.pos + base + 4 * t);
        cp_async16(&sm.meta[4 * t], b.meta + base + 4 * t);
        cp_async16(&sm.tid[4 * t], b.tid + base + 4 * t);
        if (t == 0) cp_async16(&sm.off[S5_TILE], b.cig_off + base + S5_TILE);
    } else {
        for (uint32_t r = t; r < n_tile; r += S5_THREADS) {
            sm.pos[r] = (uint32_t)b.pos[base + r]; sm.meta[r] = b.meta[base + r]; sm.tid[r] = (uint32_t)b.tid[base + r];
        }
        for (uint32_t r = t; r <= n_tile; r += S5_THREADS) sm.off[r] = b.cig_off[base + r];
    }
    if (t == 0) { sm.n_work = 0; sm.n_out = 0; }
    cp_async_commit();
    // ---- CIGAR slab of the tile.  Its address depends on cig_off; the per-tile offsets gathered by the pre-pass
    // (tile_offsets_kernel, L2-resident) let the slab request go out together with the metadata instead of one
    // DRAM round trip later.
    uint32_t lo, hi;
    if (tile_off) { lo = __ldg(tile_off + blockIdx.x); hi = __ldg(tile_off + blockIdx.x + 1); }
    else { cp_async_wait_all(); __syncthreads(); lo = sm.off[0]; hi = sm.off[n_tile]; }
    const uint32_t a0 = lo & ~3u;
    // words of the tile's slab staged in shared memory (from a0): all of it unless the tile is denser than S5_SLAB or
    // touches the ragged end of the array; alignments outside the staged window read their ops from global memory
    uint32_t n_st = 0;
    if (hi > lo) {
        const uint32_t end = min(min((hi + 3u) & ~3u, a0 + (uint32_t)S5_SLAB), b.n_ops & ~3u);
        n_st = end > a0 ? end - a0 : 0u;
        for (uint32_t v = t; v < (n_st >> 2); v += S5_THREADS) cp_async16(&sm.slab[4 * v], b.cigar + a0 + 4 * v);
    }
    cp_async_commit();
    if (tile_off) { cp_async_wait_all(); __syncthreads(); }
    cp_async_commit();
    {
        const uint4 o = *reinterpret_cast<const uint4*>(&sm.off[4 * t]);
        const uint32_t o4 = sm.off[4 * t + 4];
        const uint32_t r0 = 4 * t;
        uint32_t flags = 0;
        if (r0 + 0 < n_tile && o.y - o.x > 1u) flags |= 1u;
        if (r0 + 1 < n_tile && o.z - o.y > 1u) flags |= 2u;
        if (r0 + 2 < n_tile && o.w - o.z > 1u) flags |= 4u;
        if (r0 + 3 < n_tile && o4 - o.w > 1u) flags |= 8u;
        const uint32_t cnt = __popc(flags);
        uint32_t x = cnt;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, dlt); if ((int)lane >= dlt) x += y; }
        uint32_t wbase = 0;
        if (lane == 31 && x) wbase = atomicAdd(&sm.n_work, x);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        uint32_t p = wbase + x - cnt;
        if (flags & 1u) sm.work[p++] = (uint16_t)(r0 + 0);
        if (flags & 2u) sm.work[p++] = (uint16_t)(r0 + 1);
        if (flags & 4u) sm.work[p++] = (uint16_t)(r0 + 2);
        if (flags & 8u) sm.work[p++] = (uint16_t)(r0 + 3);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- walk the compacted alignments, one per thread, in rounds of 128
    const uint32_t n_work = (prm.debug & 1u) ? 0u : sm.n_work;
    uint32_t jstrand = 0;                                      // j1.strand == "" before an alignment's first junction
    int32_t rspan[2] = {0, 0};
    const S5Emit<Smem, S5_OUT, MOTIF, VREG, BC> emit{sm, out, cap, counters, prm.debug, &prm, &jstrand, rspan};
    auto flush = [&]() {                                        // warp 0
        const uint32_t n_st = (prm.debug & 8u) ? 0u : min(sm.n_out, (uint32_t)S5_OUT);
        __syncwarp();
        if (n_st) {
            uint32_t fb = 0;
            if (lane == 0) fb = atomicAdd(&counters[CTR_NCAND], n_st);
            fb = __shfl_sync(0xffffffffu, fb, 0);
            uint4* o = reinterpret_cast<uint4*>(out);
            for (uint32_t v = lane; v < 2 * n_st; v += 32) {
                if (fb + (v >> 1) < cap) o[2ull * fb + v] = sm.out[v];
                else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
            }
        }
        __syncwarp();
        if (lane == 0) sm.n_out = 0;
    };
    for (uint32_t w0 = 0; w0 < n_work; w0 += S5_THREADS) {
        const uint32_t w = w0 + t;
        if (w < n_work) {
            const uint32_t r = sm.work[w];
            const int32_t tid = (int32_t)sm.tid[r];
            if (tid >= 0) {
                const uint32_t o0 = sm.off[r], n = sm.off[r + 1] - o0;
                uint32_t strand = read_strand(sm.meta[r], prm.strandness);
                if (BC) strand |= (__ldg(b.bc + base + r) + 1u) << 8;      // set_junction_barcode (:362-374): one barcode per alignment
                const uint64_t read_ord = b.first_ordinal + base + r;
                if (MOTIF) jstrand = 0;
                if (VREG) {                                    // endpos = pos + reference length of the CIGAR (sam.c:327-342)
                    const bool in_smem = (o0 - a0) + n <= n_st;
                    uint32_t rl = 0;
                    for (uint32_t q = 0; q < n; ++q) {
                        const uint32_t x = in_smem ? sm.slab[o0 - a0 + q] : __ldg(b.cigar + o0 + q);
                        if ((0x18Du >> (x & 0xfu)) & 1u) rl += x >> 4;      // M, D, N, =, X consume the reference
                    }
                    // bam_endpos (sam.c:336-342): an alignment flagged BAM_FUNMAP spans one base whatever its CIGAR says
                    const bool unmapped = ((sm.meta[r] >> 16) & 4u) != 0;
                    rspan[0] = (int32_t)sm.pos[r]; rspan[1] = (int32_t)(sm.pos[r] + (unmapped ? 1u : rl));
                }
                if ((o0 - a0) + n <= n_st) walk_fast<true>(sm.slab + (o0 - a0), n, sm.pos[r], tid, strand, read_ord, emit);
                else walk_fast<false>(b.cigar + o0, n, sm.pos[r], tid, strand, read_ord, emit);
            }
        }
        __syncthreads();
        if (w0 == 0 && rg.base && !(prm.debug & 8u)) {
            // first round -> this tile's own region: plain stores by every thread, nobody waits for an atomic
            const uint32_t n_st = min(sm.n_out, (uint32_t)S5_OUT);
            uint4* o = reinterpret_cast<uint4*>(rg.base + (size_t)blockIdx.x * rg.cap);
            for (uint32_t v = t; v < 2 * n_st; v += S5_THREADS) o[v] = sm.out[v];
            if (t == 0) { rg.cnt[blockIdx.x] = n_st; if (n_st) atomicAdd(&counters[CTR_NREGION], n_st); }   // no return value used: a RED, nobody waits
            if (w0 + S5_THREADS < n_work) { __syncthreads(); if (t == 0) sm.n_out = 0; __syncthreads(); }
        } else {
            if (t < 32) flush();
            if (w0 + S5_THREADS < n_work) __syncthreads();
        }
    }
    if (n_work == 0 && rg.base && t == 0) rg.cnt[blockIdx.x] = 0;
}

// ------------------------------------------------------------------------------------------------
// cigar_scan, gather variant (variant 7; opt-in A/B build written at the end of round 1, NOT yet measured)
// ------------------------------------------------------------------------------------------------
// Why: Little's law on the round-1 numbers of the kernel above — 12 resident blocks x (8 KB columns + 2.7 KB slab) in flight
// per SM over a ~6.5 us block life = 2.9 TB/s, which is the measured rate.  The 12 comes from two limits at once
// (40 registers x 128 threads, 19.5 KB shared memory).  Only `cig_off` is needed for every alignment (to find n_cigar > 1);
// pos / meta / tid matter for the ~17 % that go on the work list.  This kernel stages only `cig_off`, the slab and the
// candidates (9.2 KB per block) and runs at <= 32 registers, so 16 blocks fit; the three columns of a work item are
// fetched by its thread straight from global memory (neighbouring work items are ~6 alignments apart, so a warp touches a
// handful of sectors per column) and that gather is issued BEFORE the wait for the slab, so it rides on the same round trip.
template <int THREADS, int SLAB, int OUTN>
struct alignas(16) S7SmemT {
    static constexpr int S5_TILE = THREADS * 4;
    uint32_t off[S5_TILE + 4];
    uint32_t slab[SLAB];
    uint4 out[OUTN * 2];
    uint32_t n_work, n_out;
    uint16_t work[S5_TILE];
};

template <int THREADS, int SLAB, int OUTN, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
cigar_scan_gather_kernel(BatchView b, ScanParams prm, Cand* __restrict__ out, uint32_t cap, uint32_t* __restrict__ counters) {
    using Smem = S7SmemT<THREADS, SLAB, OUTN>;
    constexpr int TILE = Smem::S5_TILE;
    __shared__ Smem sm;
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t base = blockIdx.x * TILE;
    const uint32_t n_tile = min((uint32_t)TILE, b.n_reads - base);
    const bool full = base + TILE + 3 <= b.n_reads;

    // ---- cig_off of the tile: global -> shared
    if (full) {
        cp_async16(&sm.off[4 * t], b.cig_off + base + 4 * t);
        if (t == 0) cp_async16(&sm.off[TILE], b.cig_off + base + TILE);
    } else {
        for (uint32_t r = t; r <= n_tile; r += THREADS) sm.off[r] = b.cig_off[base + r];
    }
    if (t == 0) { sm.n_work = 0; sm.n_out = 0; }
    cp_async_commit();
    cp_async_wait_all();
    __syncthreads();
    // ---- CIGAR slab (same window rules as cigar_scan_small_kernel)
    const uint32_t lo = sm.off[0], hi = sm.off[n_tile];
    const uint32_t a0 = lo & ~3u;
    uint32_t n_st = 0;
    if (hi > lo) {
        const uint32_t end = min(min((hi + 3u) & ~3u, a0 + (uint32_t)SLAB), b.n_ops & ~3u);
        n_st = end > a0 ? end - a0 : 0u;
        for (uint32_t v = t; v < (n_st >> 2); v += THREADS) cp_async16(&sm.slab[4 * v], b.cigar + a0 + 4 * v);
    }
    cp_async_commit();
    // ---- work list of the alignments with n_cigar > 1 (junctions_extractor.cc:379)
    {
        const uint4 o = *reinterpret_cast<const uint4*>(&sm.off[4 * t]);
        const uint32_t o4 = sm.off[4 * t + 4];
        const uint32_t r0 = 4 * t;
        uint32_t flags = 0;
        if (r0 + 0 < n_tile && o.y - o.x > 1u) flags |= 1u;
        if (r0 + 1 < n_tile && o.z - o.y > 1u) flags |= 2u;
        if (r0 + 2 < n_tile && o.w - o.z > 1u) flags |= 4u;
        if (r0 + 3 < n_tile && o4 - o.w > 1u) flags |= 8u;
        const uint32_t cnt = __popc(flags);
        uint32_t x = cnt;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, dlt); if ((int)lane >= dlt) x += y; }
        uint32_t wbase = 0;
        if (lane == 31 && x) wbase = atomicAdd(&sm.n_work, x);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        uint32_t p = wbase + x - cnt;
        if (flags & 1u) sm.work[p++] = (uint16_t)(r0 + 0);
        if (flags & 2u) sm.work[p++] = (uint16_t)(r0 + 1);
        if (flags & 4u) sm.work[p++] = (uint16_t)(r0 + 2);
        if (flags & 8u) sm.work[p++] = (uint16_t)(r0 + 3);
    }
    __syncthreads();                                            // work list complete
    const uint32_t n_work = sm.n_work;
    // ---- columns of the first round's work item, requested before the slab is waited for
    uint32_t g_r = 0, g_pos = 0, g_meta = 0;
    int32_t g_tid = -1;
    if (t < n_work) {
        g_r = sm.work[t];
        g_tid = __ldg(b.tid + base + g_r); g_pos = (uint32_t)__ldg(b.pos + base + g_r); g_meta = __ldg(b.meta + base + g_r);
    }
    cp_async_wait_all();
    __syncthreads();

    uint32_t jstrand = 0;
    int32_t rspan[2] = {0, 0};
    const S5Emit<Smem, OUTN, false, false, false> emit{sm, out, cap, counters, 0u, &prm, &jstrand, rspan};
    auto flush = [&]() {                                        // warp 0
        const uint32_t n_out = min(sm.n_out, (uint32_t)OUTN);
        __syncwarp();
        if (n_out) {
            uint32_t fb = 0;
            if (lane == 0) fb = atomicAdd(&counters[CTR_NCAND], n_out);
            fb = __shfl_sync(0xffffffffu, fb, 0);
            uint4* o = reinterpret_cast<uint4*>(out);
            for (uint32_t v = lane; v < 2 * n_out; v += 32) {
                if (fb + (v >> 1) < cap) o[2ull * fb + v] = sm.out[v];
                else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
            }
        }
        __syncwarp();
        if (lane == 0) sm.n_out = 0;
    };
    for (uint32_t w0 = 0; w0 < n_work; w0 += THREADS) {
        const uint32_t w = w0 + t;
        if (w < n_work) {
            uint32_t r = g_r, pos = g_pos, meta = g_meta;
            int32_t tid = g_tid;
            if (w0) {                                           // later rounds of a dense tile: plain loads
                r = sm.work[w];
                tid = __ldg(b.tid + base + r); pos = (uint32_t)__ldg(b.pos + base + r); meta = __ldg(b.meta + base + r);
            }
            if (tid >= 0) {
                const uint32_t o0 = sm.off[r], n = sm.off[r + 1] - o0;
                const uint32_t strand = read_strand(meta, prm.strandness);
                const uint64_t read_ord = b.first_ordinal + base + r;
                if ((o0 - a0) + n <= n_st) walk_fast<true>(sm.slab + (o0 - a0), n, pos, tid, strand, read_ord, emit);
                else walk_fast<false>(b.cigar + o0, n, pos, tid, strand, read_ord, emit);
            }
        }
        __syncthreads();
        if (t < 32) flush();
        if (w0 + THREADS < n_work) __syncthreads();
    }
}

struct WalkCand { uint32_t start, end, left, right, k; };

// Forward walk that records the first two N ops in registers (c0, c1) and counts all of them.  Predicated, no
// branches and no calls: same arithmetic as walk_fast, the first four ops are fetched by independent loads (op code
// 15 is a transparent filler), the rest in a loop.  Alignments with more than two N ops (rare) are finished by
// fused_walk_rest.  Returns the number of N ops.
template <bool FROM_SMEM>
__device__ __forceinline__ uint32_t walk_collect(const uint32_t* __restrict__ ops, uint32_t n, uint32_t pos,
                                                 WalkCand& c0, WalkCand& c1) {
    const uint32_t ANC = (1u << 0) | (1u << 7);
    const uint32_t BRK = (1u << 3) | (1u << 2) | (1u << 8) | (1u << 1) | (1u << 4);
    const uint32_t REFC = ANC | (1u << 2) | (1u << 8) | (1u << 3);
    constexpr int PRE = 4;
    uint32_t w[PRE];
#pragma unroll
    for (int i = 0; i < PRE; ++i) w[i] = (uint32_t)i < n ? (FROM_SMEM ? ops[i] : __ldg(ops + i)) : 0xfu;
    uint32_t cur = pos, run = 0, nc = 0;
    bool pending = false;
    auto step = [&](uint32_t x, uint32_t i) {
        const uint32_t op = x & 0xfu, len = x >> 4, bit = 1u << op;
        const bool brk = (bit & BRK) != 0, is_n = op == 3u;
        const bool close = pending && brk;                  // the open junction ends here: right anchor = run
        c0.right = (close && nc == 1u) ? run : c0.right;
        c1.right = (close && nc == 2u) ? run : c1.right;
        const bool open0 = is_n && nc == 0u, open1 = is_n && nc == 1u;
        const uint32_t kk = i > 0xffffu ? 0xffffu : i;
        c0.start = open0 ? cur : c0.start; c0.end = open0 ? cur + len : c0.end; c0.left = open0 ? run : c0.left; c0.k = open0 ? kk : c0.k;
        c1.start = open1 ? cur : c1.start; c1.end = open1 ? cur + len : c1.end; c1.left = open1 ? run : c1.left; c1.k = open1 ? kk : c1.k;
        nc += is_n ? 1u : 0u;
        pending = brk ? is_n : pending;
        run = brk ? 0u : run + ((bit & ANC) ? len : 0u);
        cur += (bit & REFC) ? len : 0u;
    };
#pragma unroll
    for (int i = 0; i < PRE; ++i) step(w[i], (uint32_t)i);
    for (uint32_t i = PRE; i < n; ++i) step(FROM_SMEM ? ops[i] : __ldg(ops + i), i);
    c0.right = (pending && nc == 1u) ? run : c0.right;
    c1.right = (pending && nc == 2u) ? run : c1.right;
    return nc;
}

// ------------------------------------------------------------------------------------------------
// cigar_scan, warp-pipelined persistent version (variant 8)
// ------------------------------------------------------------------------------------------------
// Round-1 finding (profiles/r1_scan_ab.md, r2_scan_ab_occupancy_variants.json): the block-per-tile kernels stream at 0.43-0.47
// of the copy peak whatever their occupancy, because every block spends most of its life with no load in flight (columns ->
// barrier -> dependent slab -> barrier -> walk -> atomic round trip -> flush).  Here nothing waits on a block:
//   * a WARP owns tiles of PW alignments (its four 512-byte column slices + its CIGAR slab) and keeps a ring of NST stages
//     in shared memory filled with 16-byte cp.async (LDGSTS); D = NST-1 tiles are always in flight per warp, so
//     warps/SM x D x ~2.7 KB stay outstanding to HBM while the warp walks the tile that has landed;
//   * the only data-dependent address — the slab, [cig_off[base], cig_off[base + PW]) — is resolved one iteration earlier by
//     two scalar loads of the NEXT tile's bounds, so the slab request leaves together with the columns;
//   * synchronisation is __syncwarp only (no __syncthreads, no named barriers, no mbarriers);
//   * candidates go straight from registers to HBM (one 32-byte sector each, neighbouring lanes neighbouring sectors) into
//     chunks of PCH slots the warp reserves with one atomic per chunk; a warp's unused tail is filled with tid = -1
//     entries, which junction_merge skips.  CTR_NCAND therefore counts reserved slots, not candidates.
// The plain mode keeps the first two N ops of an alignment in registers (walk_collect, branch-free) and stores them
// after a warp prefix sum; the third and later N ops, and the intron-motif / variant-region / barcode modes, use the generic
// per-candidate path (a shared-memory cursor into the warp's chunk).
constexpr int PW  = 128;                          // alignments per warp tile (4 per lane)
constexpr int PCH = 64;                           // candidate slots per reserved chunk

template <int SL>
struct alignas(16) PipeStage {
    uint32_t off[PW + 4];
    uint32_t pos[PW];
    uint32_t meta[PW];
    uint32_t tid[PW];
    uint32_t slab[SL];
};
constexpr int PN_CAP = 128;                       // N ops listed per detection round (one per lane per pass, four passes at most)
template <int SL, int NST, bool NLIST>
struct alignas(16) PipeWarpSmem {
    PipeStage<SL> st[NST];
    uint32_t cur, end;                            // generic path: cursor into the warp's reserved chunk
    uint32_t pad[2];
    uint8_t  work[PW];                            // generic path: alignments with more than one CIGAR op
    uint8_t  nlist[NLIST ? PN_CAP : 4];           // plain path: positions (0..127) of the N ops found in the current 128 slab words
};

__device__ __forceinline__ void store_cand(Cand* __restrict__ out, uint32_t cap, uint32_t* counters, uint32_t idx,
                                           const uint4& a, const uint4& b) {
    if (idx < cap) { uint4* o = reinterpret_cast<uint4*>(out + idx); o[0] = a; o[1] = b; }
    else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
}

// generic per-candidate emit of the pipelined kernel (any lane, any time)
template <class WS, bool MOTIF, bool VREG, bool BC>
struct PipeEmit {
    WS& ws; Cand* __restrict__ out; uint32_t cap; uint32_t* counters;
    const ScanParams* prm; uint32_t* jstrand; const int32_t* rspan;
    __device__ __forceinline__ void push(const uint4& a, const uint4& b) const {
        uint32_t i = atomicAdd(&ws.cur, 1u);
        if (i >= ws.end) i = atomicAdd(&counters[CTR_NCAND], 1u);          // chunk exhausted mid-round: one slot at a time
        store_cand(out, cap, counters, i, a, b);
    }
    __device__ __forceinline__ void operator()(uint32_t start, uint32_t end, uint32_t left, uint32_t right,
                                               uint64_t ord, int32_t tid, uint32_t strand) const {
        if (MOTIF) {                                           // set_junction_strand with a FASTA (:345-359): motif first
            const uint32_t m = motif_strand(*prm, tid, start, end, *jstrand, counters);
            if (BC) { if (m != '?') strand = (strand & ~0xffu) | m; *jstrand = strand & 0xffu; }
            else { if (m != '?') strand = m; *jstrand = strand; }
        }
        const uint4 a = make_uint4(start, end, start - left, end + right);
        const uint4 b = make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)tid, strand);
        if (!VREG) { push(a, b); return; }
        // one candidate per variant region the ALIGNMENT belongs to (tid, pos < end, endpos > beg; hts.c:1941-1963)
        const VariantRegions& vr = prm->vr;
        const int32_t rp = rspan[0], re = rspan[1];
        uint32_t lo = 0, hi = vr.n;                            // first region with (tid, beg) >= (tid, endpos)
        while (lo < hi) {
            const uint32_t mid = (lo + hi) >> 1;
            const int32_t mt = vr.tid[mid];
            if (mt < tid || (mt == tid && vr.beg[mid] < re)) lo = mid + 1; else hi = mid;
        }
        for (uint32_t i = lo; i-- > 0;) {
            if (vr.tid[i] != tid) break;
            if ((long long)vr.beg[i] + (long long)vr.max_len <= (long long)rp) break;    // no earlier region can reach pos
            if (vr.end[i] > rp) push(a, make_uint4(b.x, b.y, b.z, strand | (i + 1u) << 8));
        }
    }
};

// third and later N ops of an alignment in the plain mode (rare): plain walk from global memory, one slot each
__device__ __noinline__ void pipe_walk_rest(const uint32_t* __restrict__ ops, uint32_t n, uint32_t pos, int32_t tid, uint32_t strand,
                                            uint64_t read_ord, Cand* __restrict__ out, uint32_t cap, uint32_t* counters) {
    const uint32_t ANC = (1u << 0) | (1u << 7);
    const uint32_t BRK = (1u << 3) | (1u << 2) | (1u << 8) | (1u << 1) | (1u << 4);
    const uint32_t REFC = ANC | (1u << 2) | (1u << 8) | (1u << 3);
    uint32_t cur = pos, run = 0, nc = 0;
    bool pending = false;
    uint32_t p_start = 0, p_end = 0, p_left = 0, p_k = 0;
    auto emit = [&]() {
        if (nc <= 2u) return;
        const uint64_t ord = read_ord << 16 | p_k;
        const uint32_t g = atomicAdd(&counters[CTR_NCAND], 1u);
        store_cand(out, cap, counters, g, make_uint4(p_start, p_end, p_start - p_left, p_end + run),
                   make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)tid, strand));
    };
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t x = __ldg(ops + i), op = x & 0xfu, len = x >> 4, bit = 1u << op;
        if (bit & BRK) {
            if (pending) emit();
            pending = op == 3u;
            if (pending) { p_start = cur; p_end = cur + len; p_left = run; p_k = i > 0xffffu ? 0xffffu : i; ++nc; }
            run = 0;
        } else if (bit & ANC) {
            run += len;
        }
        if (bit & REFC) cur += len;
    }
    if (pending) emit();
}

template <int N> __device__ __forceinline__ void cp_async_wait_n() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Intron-motif / variant-region / barcode modes: processes the tile that sits in stage `st` (PW alignments from `base`) —
// work list, walk with the per-candidate emit (PipeEmit), candidate stores through the warp's shared-memory chunk cursor.
template <int SL, class WS, bool MOTIF, bool VREG, bool BC>
__device__ __forceinline__ void pipe_process_tile(WS& ws, PipeStage<SL>& st, const BatchView& b, const ScanParams& prm, uint32_t base,
                                                  uint32_t lane, uint32_t vec_end, Cand* __restrict__ out, uint32_t cap,
                                                  uint32_t* __restrict__ counters) {
    const uint32_t n_tile = min((uint32_t)PW, b.n_reads - base);
    const uint32_t lo = st.off[0], hi = st.off[n_tile], a0 = lo & ~3u;
    uint32_t n_st = 0;                                         // words of the slab staged in shared memory (from a0)
    if (hi > lo) { const uint32_t end = min(min((hi + 3u) & ~3u, a0 + (uint32_t)SL), vec_end); n_st = end > a0 ? end - a0 : 0u; }

    // ---- work list of the alignments with more than one CIGAR op (junctions_extractor.cc:379)
    uint32_t n_work;
    {
        const uint32_t lt = (1u << lane) - 1u;
        const uint4 o = *reinterpret_cast<const uint4*>(&st.off[4 * lane]);
        const uint32_t o4 = st.off[4 * lane + 4];
        const uint32_t r0 = 4 * lane;
        const bool f0 = r0 + 0 < n_tile && o.y - o.x > 1u, f1 = r0 + 1 < n_tile && o.z - o.y > 1u;
        const bool f2 = r0 + 2 < n_tile && o.w - o.z > 1u, f3 = r0 + 3 < n_tile && o4 - o.w > 1u;
        const uint32_t m0 = __ballot_sync(0xffffffffu, f0), m1 = __ballot_sync(0xffffffffu, f1);
        const uint32_t m2 = __ballot_sync(0xffffffffu, f2), m3 = __ballot_sync(0xffffffffu, f3);
        const uint32_t s1 = __popc(m0), s2 = s1 + __popc(m1), s3 = s2 + __popc(m2);
        n_work = s3 + __popc(m3);
        if (f0) ws.work[__popc(m0 & lt)] = (uint8_t)(r0 + 0);
        if (f1) ws.work[s1 + __popc(m1 & lt)] = (uint8_t)(r0 + 1);
        if (f2) ws.work[s2 + __popc(m2 & lt)] = (uint8_t)(r0 + 2);
        if (f3) ws.work[s3 + __popc(m3 & lt)] = (uint8_t)(r0 + 3);
    }
    __syncwarp();

    // ---- walk, one alignment per lane, rounds of 32
    for (uint32_t w0 = 0; w0 < n_work; w0 += 32) {
        const uint32_t w = w0 + lane;
        // make sure the round starts with a chunk that holds 64 more candidates; pad what is left of the old one
        if (ws.end - min(ws.cur, ws.end) < 64u) {
            const uint32_t pc = min(ws.cur, ws.end), pe = ws.end;
            __syncwarp();
            for (uint32_t i = pc + lane; i < pe; i += 32) store_cand(out, cap, counters, i, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0xffffffffu, 0));
            uint32_t nbase = 0;
            if (lane == 0) nbase = atomicAdd(&counters[CTR_NCAND], 2u * PCH);
            nbase = __shfl_sync(0xffffffffu, nbase, 0);
            if (lane == 0) { ws.cur = nbase; ws.end = nbase + 2u * PCH; }
            __syncwarp();
        }
        if (w < n_work) {
            const uint32_t r = ws.work[w];
            const int32_t tid = (int32_t)st.tid[r];
            if (tid >= 0) {
                const uint32_t o0 = st.off[r], n = st.off[r + 1] - o0;
                uint32_t strand = read_strand(st.meta[r], prm.strandness);
                if (BC) strand |= (__ldg(b.bc + base + r) + 1u) << 8;      // set_junction_barcode (:362-374): one barcode per alignment
                const uint64_t read_ord = b.first_ordinal + base + r;
                uint32_t jstrand = 0;                         // j1.strand == "" before an alignment's first junction
                int32_t rspan[2] = {0, 0};
                const bool in_smem = (o0 - a0) + n <= n_st;
                if (VREG) {                                    // endpos = pos + reference length of the CIGAR (sam.c:327-342)
                    uint32_t rl = 0;
                    for (uint32_t q = 0; q < n; ++q) {
                        const uint32_t x = in_smem ? st.slab[o0 - a0 + q] : __ldg(b.cigar + o0 + q);
                        if ((0x18Du >> (x & 0xfu)) & 1u) rl += x >> 4;      // M, D, N, =, X consume the reference
                    }
                    // bam_endpos (sam.c:336-342): an alignment flagged BAM_FUNMAP spans one base whatever its CIGAR says
                    const bool unmapped = ((st.meta[r] >> 16) & 4u) != 0;
                    rspan[0] = (int32_t)st.pos[r]; rspan[1] = (int32_t)(st.pos[r] + (unmapped ? 1u : rl));
                }
                const PipeEmit<WS, MOTIF, VREG, BC> emit{ws, out, cap, counters, &prm, &jstrand, rspan};
                if (in_smem) walk_fast<true>(st.slab + (o0 - a0), n, st.pos[r], tid, strand, read_ord, emit);
                else walk_fast<false>(b.cigar + o0, n, st.pos[r], tid, strand, read_ord, emit);
            }
        }
        __syncwarp();
    }
}

// Plain mode (no FASTA, no variant regions, no barcodes): OP-PARALLEL.  parse_alignment_into_junctions
// (junctions_extractor.cc:377-497) emits exactly one candidate per N op, and in closed form that candidate
// depends only on the ops of its own alignment around it.  So the tile's CIGAR slab is searched for N ops directly — four
// words per lane, ballot compaction — and each N op found gets a lane that
//   * finds its alignment r by binary search in the tile's cig_off column (shared memory),
//   * sums the reference-consuming lengths of the ops before it (start) and the M/= runs on either side (anchors),
//   * stores the candidate into the warp's chunk.
// The ~55 % of multi-op alignments that carry no N op (soft clips, indels) are never looked at beyond `(word & 15) == 3`, and
// nothing is walked op by op through a state machine.  (Round-2 measurements that led here: the per-alignment walk kernels,
// block-tiled or warp-pipelined, sit at 32-42 M warp instructions per 10 M alignments and ~55 % issue utilisation —
// instruction-bound, not memory-bound; profiles/r2_scan_*.)
template <int SL, class WS>
__device__ __forceinline__ void pipe_nops_tile(WS& ws, PipeStage<SL>& st, const BatchView& b, const ScanParams& prm, uint32_t base,
                                               uint32_t lane, uint32_t vec_end, Cand* __restrict__ out, uint32_t cap,
                                               uint32_t* __restrict__ counters, uint32_t& c_cur, uint32_t& c_end) {
    const uint32_t ANC = (1u << 0) | (1u << 7);
    const uint32_t BRK = (1u << 3) | (1u << 2) | (1u << 8) | (1u << 1) | (1u << 4);
    const uint32_t REFC = ANC | (1u << 2) | (1u << 8) | (1u << 3);
    const uint32_t n_tile = min((uint32_t)PW, b.n_reads - base);
    const uint32_t lo = st.off[0], hi = st.off[n_tile];
    if (hi <= lo) return;
    const uint32_t a0 = lo & ~3u;
    const uint32_t end_st = min(min((hi + 3u) & ~3u, a0 + (uint32_t)SL), vec_end);
    const uint32_t n_st = end_st > a0 ? end_st - a0 : 0u;     // words of the slab staged in shared memory (from a0)
    const uint32_t lt = (1u << lane) - 1u;
    auto op_at = [&](uint32_t j) -> uint32_t {                  // CIGAR word j of the batch
        return j - a0 < n_st ? st.slab[j - a0] : __ldg(b.cigar + j);
    };
    for (uint32_t v0 = a0; v0 < hi; v0 += 128) {              // 128 slab words per round, 4 per lane
        const uint32_t j0 = v0 + 4 * lane;
        uint4 w;
        if (j0 + 4 - a0 <= n_st) w = *reinterpret_cast<const uint4*>(&st.slab[j0 - a0]);
        else {                                                // outside the staged window (dense tile / ragged end of the array)
            w.x = j0 + 0 < hi ? __ldg(b.cigar + j0 + 0) : 0u; w.y = j0 + 1 < hi ? __ldg(b.cigar + j0 + 1) : 0u;
            w.z = j0 + 2 < hi ? __ldg(b.cigar + j0 + 2) : 0u; w.w = j0 + 3 < hi ? __ldg(b.cigar + j0 + 3) : 0u;
        }
        uint32_t fm = ((w.x & 0xfu) == 3u ? 1u : 0u) | ((w.y & 0xfu) == 3u ? 2u : 0u) | ((w.z & 0xfu) == 3u ? 4u : 0u) | ((w.w & 0xfu) == 3u ? 8u : 0u);
        // words in front of the tile's first op (alignment of a0) and behind its last belong to other tiles
        if (j0 < lo) fm &= 0xfu << (lo - j0);
        if (j0 + 4 > hi) fm &= j0 < hi ? 0xfu >> (j0 + 4 - hi) : 0u;
        uint32_t n_n = 0;
        for (uint32_t any = __ballot_sync(0xffffffffu, fm != 0u); any; any = __ballot_sync(0xffffffffu, fm != 0u)) {
            if (fm) { const uint32_t c = __ffs(fm) - 1; ws.nlist[n_n + __popc(any & lt)] = (uint8_t)(4 * lane + c); fm &= fm - 1; }
            n_n += __popc(any);
        }
        if (n_n == 0) continue;
        __syncwarp();
        // ---- one lane per N op (at most 128 per round: up to four passes)
        for (uint32_t w0 = 0; w0 < n_n; w0 += 32) {
            bool have = w0 + lane < n_n;
            uint32_t start = 0, end = 0, left = 0, right = 0, k = 0, r = 0;
            int32_t tid = -1;
            if (have) {
                const uint32_t J = v0 + ws.nlist[w0 + lane];
                // alignment of op J: the last r with cig_off[r] <= J (cig_off[0] = lo <= J < hi = cig_off[n_tile])
                uint32_t rl = 0, rh = n_tile;
                while (rh - rl > 1u) { const uint32_t mid = (rl + rh) >> 1; if (st.off[mid] <= J) rl = mid; else rh = mid; }
                r = rl;
                const uint32_t o0 = st.off[r], n = st.off[r + 1] - o0;
                tid = (int32_t)st.tid[r];
                have = n > 1u && tid >= 0;                    // junctions_extractor.cc:379; tid -1: no contig to name
                if (have) {
                    k = J - o0;
                    uint32_t refsum = 0;
                    bool open = true;
                    for (uint32_t i = k; i-- > 0u;) {         // ops before the N op: reference offset, left anchor
                        const uint32_t x = op_at(o0 + i), bit = 1u << (x & 0xfu), len = x >> 4;
                        refsum += (bit & REFC) ? len : 0u;
                        open = open && !(bit & BRK);
                        left += (open && (bit & ANC)) ? len : 0u;
                    }
                    for (uint32_t i = k + 1; i < n; ++i) {     // ops behind it, up to the first one that ends the exon
                        const uint32_t x = op_at(o0 + i), bit = 1u << (x & 0xfu);
                        if (bit & BRK) break;
                        right += (bit & ANC) ? x >> 4 : 0u;
                    }
                    start = st.pos[r] + refsum;
                    end = start + (op_at(J) >> 4);
                    k = k > 0xffffu ? 0xffffu : k;
                }
            }
            const uint32_t hm = __ballot_sync(0xffffffffu, have);
            const uint32_t tot = __popc(hm);
            if (tot) {
                const uint32_t rem = c_end - c_cur;
                uint32_t nbase = 0;
                if (tot > rem) {                              // the chunk runs out inside this round: the rest goes to a new one
                    if (lane == 0) nbase = atomicAdd(&counters[CTR_NCAND], (uint32_t)PCH);
                    nbase = __shfl_sync(0xffffffffu, nbase, 0);
                }
                if (have) {
                    const uint64_t ord = (b.first_ordinal + base + r) << 16 | k;
                    const uint32_t j = __popc(hm & lt);
                    store_cand(out, cap, counters, j < rem ? c_cur + j : nbase + (j - rem), make_uint4(start, end, start - left, end + right),
                               make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)tid, read_strand(st.meta[r], prm.strandness)));
                }
                if (tot > rem) { c_cur = nbase + (tot - rem); c_end = nbase + (uint32_t)PCH; } else c_cur += tot;
            }
        }
        __syncwarp();                                         // nlist is rewritten by the next round
    }
}

// The warp-pipelined kernel: cp.async (LDGSTS) into a ring of NST stages, NST - 1 tiles in flight per warp.
template <int SL, int NST, int NWARP, int MINB, bool MOTIF, bool VREG, bool BC>
__global__ void __launch_bounds__(NWARP * 32, MINB)
cigar_scan_pipe_kernel(BatchView b, ScanParams prm, Cand* __restrict__ out, uint32_t cap, uint32_t* __restrict__ counters) {
    constexpr bool GENERIC = MOTIF || VREG || BC;
    using WS = PipeWarpSmem<SL, NST, !GENERIC>;
    static_assert(NST >= 2, "the ring needs two stages");
    static_assert(SL % 128 == 0 && SL <= 512, "slab window: a multiple of 128 words, at most four 16-byte copies per lane");
    constexpr int D = NST - 1;                                // tiles in flight per warp
    constexpr int SLV = SL / 128;                             // 16-byte slab copies per lane
    extern __shared__ __align__(16) unsigned char pipe_raw[];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    WS& ws = reinterpret_cast<WS*>(pipe_raw)[warp];
    const uint32_t n_wt = (b.n_reads + PW - 1) / PW;           // warp tiles in the batch
    const uint32_t TW = gridDim.x * NWARP, gw = blockIdx.x * NWARP + warp;
    const uint32_t n_my = gw < n_wt ? (n_wt - gw + TW - 1) / TW : 0u;
    const uint32_t vec_end = b.n_ops & ~3u;
    if (GENERIC) { if (lane == 0) { ws.cur = 0; ws.end = 0; } __syncwarp(); }
    if (n_my == 0) return;

    // Slab bounds (the only data-dependent address of a tile) are fetched 16 tiles at a time, one batch ahead: lane 2j holds
    // cig_off[base] and lane 2j+1 cig_off[min(base + PW, n_reads)] of tile 16q + j, so no tile waits for them.
    auto load_bq = [&](uint32_t q) -> uint32_t {
        const uint32_t kk = q * 16u + (lane >> 1);
        if (kk >= n_my) return 0u;
        const uint32_t base = (gw + kk * TW) * PW;
        return __ldg(b.cig_off + ((lane & 1u) ? min(base + (uint32_t)PW, b.n_reads) : base));
    };
    uint32_t bq_cur = load_bq(0), bq_next = load_bq(1);
    uint32_t c_cur = 0, c_end = 0;                            // plain mode: the warp's chunk (uniform registers)

    auto issue = [&](uint32_t kk) {                           // called for kk = 0, 1, 2, ... in order, once each
        if (kk != 0u && (kk & 15u) == 0u) { bq_cur = bq_next; bq_next = load_bq((kk >> 4) + 1u); }
        const uint32_t lo = __shfl_sync(0xffffffffu, bq_cur, 2 * (kk & 15u)), hi = __shfl_sync(0xffffffffu, bq_cur, 2 * (kk & 15u) + 1);
        const uint32_t base = (gw + kk * TW) * PW;
        PipeStage<SL>& st = ws.st[kk % NST];
        if (base + PW <= b.n_reads) {                         // a full tile: its closing offset is the slab's upper bound
            cp_async16(&st.off[4 * lane], b.cig_off + base + 4 * lane);
            cp_async16(&st.pos[4 * lane], b.pos + base + 4 * lane);
            cp_async16(&st.meta[4 * lane], b.meta + base + 4 * lane);
            cp_async16(&st.tid[4 * lane], b.tid + base + 4 * lane);
            if (lane == 0) st.off[PW] = hi;
        } else {                                              // ragged tail of the batch: plain loads
            const uint32_t n_tile = b.n_reads - base;
            for (uint32_t r = lane; r < n_tile; r += 32) {
                st.pos[r] = (uint32_t)b.pos[base + r]; st.meta[r] = b.meta[base + r]; st.tid[r] = (uint32_t)b.tid[base + r];
            }
            for (uint32_t r = lane; r <= n_tile; r += 32) st.off[r] = b.cig_off[base + r];
        }
        if (hi > lo) {
            const uint32_t a0 = lo & ~3u;
            const uint32_t end = min(min((hi + 3u) & ~3u, a0 + (uint32_t)SL), vec_end);
            const uint32_t nv = end > a0 ? (end - a0) >> 2 : 0u;
            const uint32_t* src = b.cigar + a0 + 4 * lane;
#pragma unroll
            for (int j = 0; j < SLV; ++j)
                if (lane + 32u * j < nv) cp_async16(&st.slab[4 * (lane + 32 * j)], src + 128 * j);
        }
    };
#pragma unroll
    for (int j = 0; j < D; ++j) {                             // prologue: D tiles in flight
        if ((uint32_t)j < n_my) issue(j);
        cp_async_commit();
    }
    for (uint32_t k = 0; k < n_my; ++k) {
        if (k + D < n_my) issue(k + D);
        cp_async_commit();
        cp_async_wait_n<D>();                                 // this lane's copies of tile k have landed ...
        __syncwarp();                                         // ... and so have every other lane's
        const uint32_t base = (gw + k * TW) * PW;
        if (GENERIC) pipe_process_tile<SL, WS, MOTIF, VREG, BC>(ws, ws.st[k % NST], b, prm, base, lane, vec_end, out, cap, counters);
        else pipe_nops_tile<SL, WS>(ws, ws.st[k % NST], b, prm, base, lane, vec_end, out, cap, counters, c_cur, c_end);
        __syncwarp();                                         // the stage and the work list are rewritten from the next iteration on
    }
    if (!GENERIC) {
        // ---- the unused tail of the warp's last chunk: entries junction_merge skips
        for (uint32_t i = c_cur + lane; i < c_end; i += 32) store_cand(out, cap, counters, i, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0xffffffffu, 0));
    } else {
        __syncwarp();
        const uint32_t pc = min(ws.cur, ws.end), pe = ws.end;
        for (uint32_t i = pc + lane; i < pe; i += 32) store_cand(out, cap, counters, i, make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0xffffffffu, 0));
    }
}

// pre-pass: tile_off[t] = cig_off[min(t * S5_TILE, n_reads)] (one 4-byte load per tile; the result stays in L2)
__global__ void tile_offsets_kernel(const uint32_t* __restrict__ cig_off, uint32_t n_reads, uint32_t n_tiles, uint32_t tile, uint32_t* __restrict__ tile_off) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t <= n_tiles) tile_off[t] = cig_off[min(t * tile, n_reads)];
}
uint32_t cigar_scan_tiles(uint32_t n_reads) { return (n_reads + 255) / 256; }   // upper bound over the tile sizes in use
static int scan_variant() {
    static int variant = -1;
    if (variant < 0) { const char* v = getenv("RTJX_SCAN_VARIANT"); variant = v ? atoi(v) : 5; if (variant != 1 && variant != 4) variant = 5; }
    return variant;
}
static int scan_cfg() {
    static int cfg = -1;
    if (cfg < 0) { const char* v = getenv("RTJX_SCAN_CFG"); cfg = v ? atoi(v) : 0; }
    return cfg;
}
void cigar_scan_region_layout(uint32_t n_reads, uint32_t* n_regions, uint32_t* cap) {
    static int use = -1;
    if (use < 0) { const char* v = getenv("RTJX_SCAN_REGIONS"); use = v ? atoi(v) : 0; }   // measured: no gain (A/B option)
    if (use && scan_variant() == 5 && scan_cfg() == 0) { *n_regions = (n_reads + 511) / 512; *cap = 192; }
    else { *n_regions = 0; *cap = 0; }
}

// Launch of the warp-pipelined kernel: persistent grid of BPS blocks per SM, NWARP warps each.
template <int SL, int NST, int NWARP, int BPS, bool MOTIF, bool VREG, bool BC>
static void launch_pipe_mode(const BatchView& b, const ScanParams& p, Cand* cands, uint32_t cand_cap, uint32_t* d_counters, cudaStream_t stream) {
    constexpr size_t smem = (size_t)NWARP * sizeof(PipeWarpSmem<SL, NST, !(MOTIF || VREG || BC)>);
    static_assert((smem + 1024) * BPS <= 228u * 1024u, "pipelined scan: shared memory of the resident blocks exceeds an SM");
    auto kern = cigar_scan_pipe_kernel<SL, NST, NWARP, BPS, MOTIF, VREG, BC>;
    static bool once = false;
    if (!once) {
        once = true;
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (getenv("RTJX_TRACE")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NWARP * 32, smem);
            fprintf(stderr, "[rtjx] cigar_scan_pipe<SL %d, NST %d, %d warps, %d blocks/SM asked>: %zu B shared memory per block, %d blocks/SM resident\n",
                    SL, NST, NWARP, BPS, smem, nb);
        }
    }
    const uint32_t n_wt = (b.n_reads + PW - 1) / PW;
    const uint32_t grid = max(1u, min((n_wt + NWARP - 1) / NWARP, (uint32_t)(num_sms() * BPS)));
    kern<<<grid, NWARP * 32, smem, stream>>>(b, p, cands, cand_cap, d_counters);
}
template <int SL, int NST, int NWARP, int BPS, bool SPECIAL = false>
static void launch_pipe(const BatchView& b, const ScanParams& p, Cand* cands, uint32_t cand_cap, uint32_t* d_counters, cudaStream_t stream) {
    if constexpr (SPECIAL) {                                  // intron-motif / variant-region / barcode modes
        if (b.bc && p.genome) launch_pipe_mode<SL, NST, NWARP, BPS, true, false, true>(b, p, cands, cand_cap, d_counters, stream);
        else if (b.bc) launch_pipe_mode<SL, NST, NWARP, BPS, false, false, true>(b, p, cands, cand_cap, d_counters, stream);
        else if (p.vr.n && p.genome) launch_pipe_mode<SL, NST, NWARP, BPS, true, true, false>(b, p, cands, cand_cap, d_counters, stream);
        else if (p.vr.n) launch_pipe_mode<SL, NST, NWARP, BPS, false, true, false>(b, p, cands, cand_cap, d_counters, stream);
        else launch_pipe_mode<SL, NST, NWARP, BPS, true, false, false>(b, p, cands, cand_cap, d_counters, stream);
    } else {
        launch_pipe_mode<SL, NST, NWARP, BPS, false, false, false>(b, p, cands, cand_cap, d_counters, stream);
    }
}
// Candidate slots cigar_scan may reserve beyond the number of N ops of a batch (every warp's last chunk is partly padding).
uint32_t cigar_scan_cand_slack() { return (uint32_t)num_sms() * 64u * 2u * (uint32_t)PCH + 1024u; }

void launch_cigar_scan(const BatchView& b, const ScanParams& p, Cand* cands, uint32_t cand_cap,
                       uint32_t* d_counters, uint32_t* tile_off_scratch, const CandRegions& regions, cudaStream_t stream) {
    if (b.n_reads == 0) return;
    const uintptr_t align = reinterpret_cast<uintptr_t>(b.tid) | reinterpret_cast<uintptr_t>(b.pos) |
                            reinterpret_cast<uintptr_t>(b.meta) | reinterpret_cast<uintptr_t>(b.cig_off) |
                            reinterpret_cast<uintptr_t>(b.cigar);
    if ((align & 15u) == 0 && p.variant == 8) {
        const bool special = p.genome || p.vr.n || b.bc;
        if (special) { launch_pipe<384, 3, 8, 2, true>(b, p, cands, cand_cap, d_counters, stream); return; }
        switch (p.cfg) {                   // A/B configurations; 0 is the production one.  Per warp: NST stages of 3.6 KB (SL 384) / 3.1 KB (SL 256)
        case 1: launch_pipe<384, 2, 8, 3>(b, p, cands, cand_cap, d_counters, stream); break;      // 24 warps/SM, 1 tile in flight each
        case 2: launch_pipe<384, 3, 4, 4>(b, p, cands, cand_cap, d_counters, stream); break;      // 16 warps/SM, 2 in flight
        case 3: launch_pipe<256, 2, 8, 4>(b, p, cands, cand_cap, d_counters, stream); break;      // 32 warps/SM, 1 in flight (64 registers)
        case 4: launch_pipe<384, 4, 4, 3>(b, p, cands, cand_cap, d_counters, stream); break;      // 12 warps/SM, 3 in flight
        case 5: launch_pipe<256, 3, 4, 5>(b, p, cands, cand_cap, d_counters, stream); break;      // 20 warps/SM, 2 in flight
        case 6: launch_pipe<384, 2, 8, 1>(b, p, cands, cand_cap, d_counters, stream); break;      // 8 warps/SM (latency probe)
        case 7: launch_pipe<256, 2, 4, 7>(b, p, cands, cand_cap, d_counters, stream); break;      // 28 warps/SM, 1 in flight
        default: launch_pipe<384, 2, 8, 3>(b, p, cands, cand_cap, d_counters, stream); break;
        }
        return;
    }
    const int variant = (p.genome || p.vr.n || b.bc) ? 5 : ((p.variant == 1 || p.variant == 4 || p.variant == 7) ? p.variant : scan_variant());   // only variant 5 knows the intron-motif and variant-region modes
    if ((align & 15u) == 0 && variant == 5) {
        static int prepass = -1;
        const int cfg = b.bc ? 0 : (p.variant == 5 && p.cfg ? p.cfg : scan_cfg());
        if (prepass < 0) { const char* v = getenv("RTJX_SCAN_PREPASS"); prepass = v ? atoi(v) : 0; }   // measured: no gain on B200
        const uint32_t threads = cfg == 2 ? 64u : cfg == 3 ? 256u : 128u;
        const uint32_t tile = threads * 4, tiles = (b.n_reads + tile - 1) / tile;
        const uint32_t* toff = nullptr;
        if (prepass && tile_off_scratch && tiles >= 64) {
            tile_offsets_kernel<<<(tiles + 1 + 255) / 256, 256, 0, stream>>>(b.cig_off, b.n_reads, tiles, tile, tile_off_scratch);
            toff = tile_off_scratch;
        }
        CandRegions none{nullptr, nullptr, 0, 0};
        switch (cfg) {          // A/B configurations (RTJX_SCAN_CFG); 0 is the production one
        case 1: cigar_scan_small_kernel<128, 768, 128><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        case 2: cigar_scan_small_kernel<64, 512, 96><<<tiles, 64, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        case 3: cigar_scan_small_kernel<256, 2048, 384><<<tiles, 256, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        case 4: cigar_scan_small_kernel<128, 1024, 96><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        case 5: cigar_scan_small_kernel<128, 2048, 192><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        case 6: cigar_scan_small_kernel<128, 2048, 160><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        case 7: cigar_scan_small_kernel<128, 1536, 160><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        // Occupancy experiments for round 2 (not yet measured).  Little's law on the round-1 numbers: 12 resident blocks x
        // (8 KB columns + 2.7 KB slab) in flight per SM over a ~6.5 us block life = 2.9 TB/s, which IS the measured rate; the
        // 12 comes from both limits at once (40 registers x 128 threads -> 12.8 blocks, 19.5 KB shared -> 11.6).  These
        // configurations lift both: <= 32 registers (launch bound 16) and <= 14 KB shared memory per block.
        case 8: cigar_scan_small_kernel<128, 768, 64, false, false, false, 16><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        case 9: cigar_scan_small_kernel<128, 768, 96, false, false, false, 14><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;
        case 10: cigar_scan_small_kernel<128, 1024, 192, false, false, false, 16><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none); break;   // registers only
        default:
            if (b.bc && p.genome) cigar_scan_small_kernel<128, 1024, 192, true, false, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none);
            else if (b.bc) cigar_scan_small_kernel<128, 1024, 192, false, false, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none);
            else if (p.vr.n && p.genome) cigar_scan_small_kernel<128, 1024, 192, true, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none);
            else if (p.vr.n) cigar_scan_small_kernel<128, 1024, 192, false, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, none);
            else if (p.genome) cigar_scan_small_kernel<128, 1024, 192, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, regions);
            else cigar_scan_small_kernel<128, 1024, 192><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters, toff, regions);
            break;
        }
        return;
    }
    if ((align & 15u) == 0 && variant == 7) {                  // gather variant, A/B configurations via scan_cfg
        const uint32_t tiles = (b.n_reads + 511u) / 512u;
        switch (p.variant == 7 ? p.cfg : scan_cfg()) {
        case 1: cigar_scan_gather_kernel<128, 1024, 96, 16><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
        case 2: cigar_scan_gather_kernel<128, 768, 96, 1><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
        case 3: cigar_scan_gather_kernel<128, 1024, 192, 16><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
        case 4: cigar_scan_gather_kernel<128, 768, 96, 14><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
        default: cigar_scan_gather_kernel<128, 768, 96, 16><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
        }
        return;
    }
    if ((align & 15u) == 0 && variant == 4) {
        static bool attr4 = false;
        if (!attr4) { cudaFuncSetAttribute(cigar_scan_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S4Smem)); attr4 = true; }
        const uint32_t tiles = (b.n_reads + S4_TILE - 1) / S4_TILE;
        cigar_scan_ws_kernel<<<min(tiles, (uint32_t)(2 * num_sms())), S4_THREADS, sizeof(S4Smem), stream>>>(b, p, cands, cand_cap, d_counters);
        return;
    }
    uint32_t grid = (b.n_reads + SCAN_TILE - 1) / SCAN_TILE;
    cigar_scan_kernel<<<grid, SCAN_THREADS, 0, stream>>>(b, p, cands, cand_cap, d_counters);
}

// ------------------------------------------------------------------------------------------------
// device-wide junction table
// ------------------------------------------------------------------------------------------------
// Upsert of an (already aggregated) partial reduction into the global table.  The thread that wins
// the 128-bit CAS on an empty slot also records the slot index in `slot_list` (position = running
// count of distinct junctions), so finalize and clear touch only occupied slots.  Returns false
// if the table has no free slot on the probe path (caller spills).
__device__ __forceinline__ bool table_upsert(const TableRef& tb, K128 key, uint32_t count, uint32_t nts, uint32_t te,
                                             uint32_t lr, unsigned long long nfirst, unsigned long long last,
                                             uint32_t* counters) {
    const uint32_t mask = tb.mask;
    uint32_t s = mix_key(key.lo, key.hi) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe, s = (s + 1) & mask) {
        Slot* sl = tb.slots + s;
        K128 cur = ld128_relaxed(sl);
        if (cur.lo == 0ull && cur.hi == 0ull) {
            cur = cas128(sl, K128{0ull, 0ull}, key);
            if (cur.lo == 0ull && cur.hi == 0ull) {
                const uint32_t idx = atomicAdd(&counters[CTR_NUNIQUE], 1u);
                if (idx < tb.list_cap) tb.slot_list[idx] = s;
                cur = key;
            }
        }
        if (cur.lo == key.lo && cur.hi == key.hi) {
            atomicAdd(&sl->count, count);
            atomicMax(&sl->nts, nts);
            atomicMax(&sl->te, te);
            if (lr) atomicOr(&sl->lr, lr);
            atomicMax(&sl->nfirst, nfirst);
            if (last) atomicMax(&sl->last, last);
            return true;
        }
    }
    return false;
}

__device__ __forceinline__ void spill_entry(Slot* __restrict__ spill, uint32_t spill_cap, uint32_t* counters, K128 key,
                                            uint32_t count, uint32_t nts, uint32_t te, uint32_t lr,
                                            unsigned long long nfirst, unsigned long long last) {
    uint32_t i = atomicAdd(&counters[CTR_NSPILL], 1u);
    if (i < spill_cap) {
        Slot s;
        s.klo = key.lo; s.khi = key.hi; s.count = count; s.nts = nts; s.te = te; s.lr = lr;
        s.nfirst = nfirst; s.last = last;
        spill[i] = s;
    }
}

// ------------------------------------------------------------------------------------------------
// junction_merge
// ------------------------------------------------------------------------------------------------
// Persistent grid; each block takes tiles of MERGE_TILE candidates.  A BAM is coordinate sorted, so
// neighbouring candidates repeat the same few junctions (a hot junction has 1e5-1e6 supporting
// reads): the tile is first reduced in a shared-memory hash keyed on a 62-bit (start, intron
// length, proxy) word — valid while every candidate of the tile is on the tile's first contig — and
// only one upsert per distinct junction per tile reaches the L2 atomics.
constexpr int MERGE_THREADS = 256;
constexpr int MERGE_CPT     = 8;
constexpr int MERGE_TILE    = MERGE_THREADS * MERGE_CPT;     // 2048 candidates
constexpr int MERGE_SLOTS   = 2048;                          // shared-memory hash slots (power of 2)
constexpr int MERGE_PROBES  = 32;
constexpr unsigned long long SKEY_EMPTY = ~0ull;
constexpr int MERGE_PAL = 4;                               // (contig, region) pairs per chunk in the shared-memory table's 62-bit key
constexpr int MERGE_RGROUP = 32;                            // scan tiles (candidate regions) per merge tile

struct MergeSmem {
    unsigned long long key[MERGE_SLOTS];
    unsigned long long nfirst[MERGE_SLOTS];
    unsigned long long last[MERGE_SLOTS];
    uint32_t count[MERGE_SLOTS], nts[MERGE_SLOTS], te[MERGE_SLOTS], lr[MERGE_SLOTS];
    unsigned long long pal[MERGE_PAL];           // region << 32 | tid of the (contig, region) pairs the chunk's shared-memory table knows
    uint32_t rpre[MERGE_RGROUP + 1];             // prefix of the region counts of a region tile
};

__global__ void __launch_bounds__(MERGE_THREADS, 2)
junction_merge_kernel(const Cand* __restrict__ cands, const uint32_t* __restrict__ d_n_cand, uint32_t n_bound, CandRegions rg,
                      ScanParams prm, TableRef tb, Slot* __restrict__ spill, uint32_t spill_cap,
                      uint32_t* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MergeSmem& sm = *reinterpret_cast<MergeSmem*>(smem_raw);
    const uint32_t t = threadIdx.x;
    const uint32_t n = d_n_cand ? min(*d_n_cand, n_bound) : n_bound;
    const uint32_t n_otiles = (n + MERGE_TILE - 1) / MERGE_TILE;                                   // dense overflow list
    const uint32_t n_rtiles = rg.base ? (rg.n_regions + MERGE_RGROUP - 1) / MERGE_RGROUP : 0u;     // per-tile regions
    uint32_t n_valid = 0;                                       // candidates seen (tid >= 0: not the padding of the scan's chunks)

    for (uint32_t tile = blockIdx.x; tile < n_rtiles + n_otiles; tile += gridDim.x) {
        // ---- where this tile's candidates are
        const bool is_region = tile < n_rtiles;
        uint32_t total, g0 = 0;
        if (is_region) {
            g0 = tile * MERGE_RGROUP;
            const uint32_t ng = min((uint32_t)MERGE_RGROUP, rg.n_regions - g0);
            if (t < 32) {                                       // warp 0: inclusive scan of the region counts
                uint32_t c = t < ng ? min(rg.cnt[g0 + t], rg.cap) : 0u, x = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, d); if ((int)t >= d) x += y; }
                sm.rpre[t + 1] = x;
                if (t == 0) sm.rpre[0] = 0;
            }
            __syncthreads();
            total = sm.rpre[MERGE_RGROUP];
        } else {
            const uint32_t tb0 = (tile - n_rtiles) * MERGE_TILE;
            total = min((uint32_t)MERGE_TILE, n - tb0);
        }
        auto cand_ptr = [&](uint32_t i) -> const uint4* {       // i-th candidate of the tile
            if (!is_region) return reinterpret_cast<const uint4*>(cands + (size_t)(tile - n_rtiles) * MERGE_TILE + i);
            uint32_t lo = 0, hi = MERGE_RGROUP;                 // last g with rpre[g] <= i
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (sm.rpre[mid] <= i) lo = mid; else hi = mid; }
            return reinterpret_cast<const uint4*>(rg.base + (size_t)(g0 + lo) * rg.cap + (i - sm.rpre[lo]));
        };

        for (uint32_t c0 = 0; c0 < total; c0 += MERGE_TILE) {   // a region tile may hold more than MERGE_TILE candidates
            const uint32_t m = min((uint32_t)MERGE_TILE, total - c0);
            for (uint32_t s = t; s < MERGE_SLOTS; s += MERGE_THREADS) {
                sm.key[s] = SKEY_EMPTY; sm.nfirst[s] = 0ull; sm.last[s] = 0ull;
                sm.count[s] = 0u; sm.nts[s] = 0u; sm.te[s] = 0u; sm.lr[s] = 0u;
            }
            if (t < MERGE_PAL) sm.pal[t] = SKEY_EMPTY;
            // all loads of the chunk first (two 128-bit loads per candidate)
            uint4 ca[MERGE_CPT], cb[MERGE_CPT];
#pragma unroll
            for (int j = 0; j < MERGE_CPT; ++j) {
                const uint32_t i = t + j * MERGE_THREADS;
                if (i < m) {
                    const uint4* p = cand_ptr(c0 + i);
                    ca[j] = ldg_stream_u4(p);
                    cb[j] = ldg_stream_u4(p + 1);
                } else {
                    ca[j] = make_uint4(0, 0, 0, 0);
                    cb[j] = make_uint4(0, 0, 0xffffffffu, 0);      // tid = -1: skipped
                }
            }
            __syncthreads();
#pragma unroll
            for (int j = 0; j < MERGE_CPT; ++j) {
                const uint32_t start = ca[j].x, end = ca[j].y, ts = ca[j].z, te = ca[j].w;
                const int32_t tid = (int32_t)cb[j].z;
                if (tid < 0) continue;
                ++n_valid;
                const uint32_t ilen = end - start;                                   // uint32, :161-162
                if (ilen < prm.min_intron || ilen > prm.max_intron) continue;         // junction_qc
                const uint32_t lr = ((start - ts) >= prm.min_anchor ? 1u : 0u) | ((te - end) >= prm.min_anchor ? 2u : 0u);
                const uint32_t sc = cb[j].w & 0xffu, vreg = cb[j].w >> 8;           // vreg = variant region + 1 (0 outside that mode)
                const uint32_t proxy = sc == '+' ? 0u : (sc == '-' ? 1u : 2u);         // :186-193
                const unsigned long long ord = (unsigned long long)cb[j].y << 32 | cb[j].x;
                const unsigned long long nfirst = ~ord;
                const unsigned long long last = proxy == 2u ? ((ord >> 16) << 8 | sc) : 0ull;
                bool done = false;
                // (contig, region) -> palette index: the first MERGE_PAL distinct pairs of the chunk (a BAM is sorted: nearly always 1-2)
                uint32_t pi = MERGE_PAL;
                {
                    const unsigned long long ck = (unsigned long long)vreg << 32 | (uint32_t)tid;
#pragma unroll
                    for (uint32_t q = 0; q < (uint32_t)MERGE_PAL; ++q) {
                        unsigned long long cur = sm.pal[q];
                        if (cur == SKEY_EMPTY) cur = atomicCAS(&sm.pal[q], SKEY_EMPTY, ck);
                        if (cur == SKEY_EMPTY || cur == ck) { pi = q; break; }
                    }
                }
                if (pi < (uint32_t)MERGE_PAL && ilen < (1u << 26)) {
                    const unsigned long long k = (unsigned long long)start << 30 | (unsigned long long)ilen << 4 | pi << 2 | proxy;
                    uint32_t s = mix_key(k, 0ull) & (MERGE_SLOTS - 1);
                    for (int probe = 0; probe < MERGE_PROBES; ++probe, s = (s + 1) & (MERGE_SLOTS - 1)) {
                        unsigned long long cur = sm.key[s];
                        if (cur == SKEY_EMPTY) cur = atomicCAS(&sm.key[s], SKEY_EMPTY, k);
                        if (cur == SKEY_EMPTY || cur == k) {
                            atomicAdd(&sm.count[s], 1u);
                            atomicMax(&sm.nts[s], ~ts);
                            atomicMax(&sm.te[s], te);
                            if (lr) atomicOr(&sm.lr[s], lr);
                            atomicMax(&sm.nfirst[s], nfirst);
                            if (last) atomicMax(&sm.last[s], last);
                            done = true;
                            break;
                        }
                    }
                }
                if (!done) {
                    K128 key{(unsigned long long)start << 32 | end, (unsigned long long)vreg << 34 | ((unsigned long long)(uint32_t)(tid + 1)) << 2 | proxy};
                    if (!table_upsert(tb, key, 1u, ~ts, te, lr, nfirst, last, counters))
                        spill_entry(spill, spill_cap, counters, key, 1u, ~ts, te, lr, nfirst, last);
                }
            }
            __syncthreads();
            // one global upsert per distinct junction of the chunk
            for (uint32_t s = t; s < MERGE_SLOTS; s += MERGE_THREADS) {
                const unsigned long long k = sm.key[s];
                if (k == SKEY_EMPTY) continue;
                const uint32_t start = (uint32_t)(k >> 30), ilen = (uint32_t)(k >> 4) & 0x03ffffffu, proxy = (uint32_t)k & 3u;
                const unsigned long long ck = sm.pal[((uint32_t)k >> 2) & 3u];
                K128 key{(unsigned long long)start << 32 | (uint32_t)(start + ilen),
                         (ck >> 32) << 34 | ((unsigned long long)((uint32_t)ck + 1u)) << 2 | proxy};
                if (!table_upsert(tb, key, sm.count[s], sm.nts[s], sm.te[s], sm.lr[s], sm.nfirst[s], sm.last[s], counters))
                    spill_entry(spill, spill_cap, counters, key, sm.count[s], sm.nts[s], sm.te[s], sm.lr[s], sm.nfirst[s], sm.last[s]);
            }
            __syncthreads();
        }
        __syncthreads();            // rpre is rewritten by the next region tile
    }
    n_valid = __reduce_add_sync(0xffffffffu, n_valid);
    if ((t & 31u) == 0 && n_valid) atomicAdd(reinterpret_cast<unsigned long long*>(counters + CTR_TOTAL_CAND64), (unsigned long long)n_valid);
}

static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

void launch_junction_merge(const Cand* cands, const uint32_t* d_n_cand, uint32_t n_cand_bound, const CandRegions& regions,
                           const ScanParams& p, const TableRef& tb, Slot* spill_slots, uint32_t spill_cap, uint32_t* d_counters,
                           cudaStream_t stream) {
    if (n_cand_bound == 0 && !regions.base) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(junction_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem));
        attr_set = true;
    }
    uint32_t tiles = (n_cand_bound + MERGE_TILE - 1) / MERGE_TILE + (regions.base ? (regions.n_regions + MERGE_RGROUP - 1) / MERGE_RGROUP : 0u);
    uint32_t grid = max(1u, min(tiles, (uint32_t)(2 * num_sms())));
    junction_merge_kernel<<<grid, MERGE_THREADS, sizeof(MergeSmem), stream>>>(
        cands, d_n_cand, n_cand_bound, regions, p, tb, spill_slots, spill_cap, d_counters);
}

// ------------------------------------------------------------------------------------------------
// cigar_scan, fused version (variant 6; opt-in: rtjx_params.scan_variant = 6)
// ------------------------------------------------------------------------------------------------
// parse_alignment_into_junctions (:377-497) + set_junction_strand (:345-359) + junction_qc (:160-170) +
// add_junction (:174-235) in ONE kernel: no candidate list in HBM, no second kernel.
// Measured on the 10M-read C2 batch: 117 us for scan + merge in one launch against 71 + 58 us for the default
// two-kernel path (variant 5); variant 5 stays the default because its scan kernel alone is the path's roofline line.
//
// Shape, as measured on B200 (tools/ab_scan.py):
//  * one block = one tile of 4*THREADS consecutive alignments, many small blocks per SM whose phases the block
//    scheduler overlaps (persistent rings with deeper prefetch were measured and are slower on this path);
//  * tiles are REGISTER-staged: 128-bit ld.global of the four metadata columns -> st.shared, then (its address is
//    the only data-dependent one) the tile's CIGAR slab the same way.  cp.async/LDGSTS and bulk copies top out at
//    ~0.66 of the copy peak for this tile pattern, plain loads reach 0.76 (the practical ceiling for 212 MB);
//  * while the slab is in flight the n_cigar > 1 alignments are compacted (ballot + warp prefix) into a work list;
//  * one thread per compacted alignment walks the CIGAR (closed form, SURVEY App. A.2, first four ops loaded
//    independently, up to two candidates kept in registers) and applies QC;
//  * the lanes of a warp that hold the same junction are combined by a leader-election loop (ballot + redux; a
//    BAM is coordinate sorted, so a warp usually holds 1-5 distinct junctions), the leaders accumulate into the
//    block's 128-slot shared-memory table (32-bit tile-local ordinals);
//  * one upsert per distinct junction of the tile goes to the device-wide table (128-bit CAS claim + REDs).
constexpr int S6_HS = 128;                                   // shared-memory table slots per block
constexpr int S6_PROBES = 16;
constexpr int S6_SV = 4;                                     // slab vectors (16 B) staged per thread: 2048 words cover a fully spliced tile
template <int THREADS>
struct alignas(16) S6Smem {
    uint32_t off[THREADS * 4 + 4];
    uint32_t pos[THREADS * 4];
    uint32_t meta[THREADS * 4];
    uint32_t tid[THREADS * 4];
    uint32_t slab[THREADS * 4 * S6_SV];
    unsigned long long hkey[S6_HS];
    uint32_t hval[6][S6_HS];                                 // count, ~thick_start, thick_end, lr, ~first(local), last(local)
    uint16_t work[THREADS * 4];
    uint32_t n_work, sink;
};

struct FusedCtx { TableRef tb; Slot* spill; uint32_t spill_cap; uint32_t* counters; };

// One (already combined) update of the device-wide table; tile-local 32-bit ordinals become global here.
// Kept out of line: it has several call sites and the scan kernel's hot path has to stay small.
__device__ __noinline__ void fused_global_upsert(const FusedCtx& cx, uint64_t ord0, uint32_t start, uint32_t end, int32_t tid,
                                                 uint32_t proxy, uint32_t count, uint32_t nts, uint32_t te, uint32_t lr,
                                                 uint32_t nfirst_l, uint32_t last_l) {
    const uint32_t lord = ~nfirst_l;
    const unsigned long long nfirst = ~(((ord0 + (lord >> 16)) << 16) | (lord & 0xffffu));
    const unsigned long long last = last_l ? ((ord0 + (last_l >> 8)) << 8 | (last_l & 0xffu)) : 0ull;
    const K128 key{(unsigned long long)start << 32 | end, ((unsigned long long)(uint32_t)(tid + 1)) << 2 | proxy};
    if (!table_upsert(cx.tb, key, count, nts, te, lr, nfirst, last, cx.counters))
    spill_entry(cx.spill, cx.spill_cap, cx.counters, key, count, nts, te, lr, nfirst, last);
}

// The third and later N ops of one alignment (long-read style CIGARs; rare): plain walk from global memory, QC,
// one global upsert each.  Out of line so that the hot walk carries no call.
__device__ __noinline__ void fused_walk_rest(const FusedCtx& cx, const ScanParams& prm, uint64_t ord0, const uint32_t* __restrict__ ops,
                                             uint32_t n, uint32_t pos, uint32_t r, int32_t tid, uint32_t sc) {
    const uint32_t ANC = (1u << 0) | (1u << 7);
    const uint32_t BRK = (1u << 3) | (1u << 2) | (1u << 8) | (1u << 1) | (1u << 4);
    const uint32_t REFC = ANC | (1u << 2) | (1u << 8) | (1u << 3);
    const uint32_t proxy = sc == '+' ? 0u : (sc == '-' ? 1u : 2u);
    uint32_t cur = pos, run = 0, nc = 0;
    bool pending = false;
    WalkCand c{0, 0, 0, 0, 0};
    auto emit = [&]() {
        const uint32_t ilen = c.end - c.start;
        if (nc <= 2u || ilen < prm.min_intron || ilen > prm.max_intron) return;
        fused_global_upsert(cx, ord0, c.start, c.end, tid, proxy, 1u, ~(c.start - c.left), c.end + c.right,
                            (c.left >= prm.min_anchor ? 1u : 0u) | (c.right >= prm.min_anchor ? 2u : 0u),
                            ~(r << 16 | c.k), proxy == 2u ? (r << 8 | sc) : 0u);
    };
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t x = __ldg(ops + i), op = x & 0xfu, len = x >> 4, bit = 1u << op;
        if (bit & BRK) {
            if (pending) { c.right = run; emit(); }
            pending = op == 3u;
            if (pending) { c.start = cur; c.end = cur + len; c.left = run; c.k = i > 0xffffu ? 0xffffu : i; ++nc; }
            run = 0;
        } else if (bit & ANC) {
            run += len;
        }
        if (bit & REFC) cur += len;
    }
    if (pending) { c.right = run; emit(); }
}

template <int THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
cigar_scan_fused_kernel(BatchView b, ScanParams prm, TableRef tb, Slot* __restrict__ spill, uint32_t spill_cap,
                        uint32_t* __restrict__ counters) {
    using Smem = S6Smem<THREADS>;
    constexpr uint32_t TILE = THREADS * 4, SLAB = THREADS * 4 * S6_SV;
    __shared__ Smem sm;
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t base = blockIdx.x * TILE;
    const uint32_t n_tile = min(TILE, b.n_reads - base);
    const uint32_t vec_end = b.n_ops & ~3u;

    // ---- phase 1: the four metadata columns, 128-bit loads -> registers -> shared memory
    if (base + TILE + 3 <= b.n_reads) {
        const uint4 o = ldg_stream_u4(reinterpret_cast<const uint4*>(b.cig_off + base) + t);
        const uint4 p = ldg_stream_u4(reinterpret_cast<const uint4*>(b.pos + base) + t);
        const uint4 m = ldg_stream_u4(reinterpret_cast<const uint4*>(b.meta + base) + t);
        const uint4 d = ldg_stream_u4(reinterpret_cast<const uint4*>(b.tid + base) + t);
        if (t == 0) sm.off[TILE] = __ldg(b.cig_off + base + TILE);
        *reinterpret_cast<uint4*>(&sm.off[4 * t]) = o;
        *reinterpret_cast<uint4*>(&sm.pos[4 * t]) = p;
        *reinterpret_cast<uint4*>(&sm.meta[4 * t]) = m;
        *reinterpret_cast<uint4*>(&sm.tid[4 * t]) = d;
    } else {                                                 // ragged tail of the batch
        for (uint32_t r = t; r < n_tile; r += THREADS) {
            sm.pos[r] = (uint32_t)b.pos[base + r]; sm.meta[r] = b.meta[base + r]; sm.tid[r] = (uint32_t)b.tid[base + r];
        }
        for (uint32_t r = t; r <= n_tile; r += THREADS) sm.off[r] = b.cig_off[base + r];
    }
    for (uint32_t s = t; s < (uint32_t)S6_HS; s += THREADS) {
        sm.hkey[s] = SKEY_EMPTY;
#pragma unroll
        for (int f = 0; f < 6; ++f) sm.hval[f][s] = 0u;
    }
    if (t == 0) { sm.n_work = 0; sm.sink = 0; }
    __syncthreads();
    if (prm.debug & 16u) return;

    // ---- phase 2: the tile's CIGAR slab [lo, hi) (the only data-dependent address of the path); at most SLAB words
    // are staged, the rest of a dense tile is warmed in L2 and read from there
    const uint32_t lo = sm.off[0], hi = sm.off[n_tile], a0 = lo & ~3u;
    uint32_t n_st = 0;
    if (hi > lo) {
        const uint32_t end = min(min((hi + 3u) & ~3u, a0 + SLAB), vec_end);
        n_st = end > a0 ? end - a0 : 0u;
    }
    uint4 sv[S6_SV];
#pragma unroll
    for (int j = 0; j < S6_SV; ++j) {
        const uint32_t v = t + j * THREADS;
        if (4u * v < n_st) sv[j] = ldg_stream_u4(reinterpret_cast<const uint4*>(b.cigar + a0) + v);
    }
    if (t == 32 % THREADS && hi > lo) {
        const uint32_t end = min((hi + 3u) & ~3u, vec_end);
        if (end > a0 + n_st) bulk_prefetch_l2(b.cigar + a0 + n_st, (end - a0 - n_st) * 4u);
    }
    // ---- while the slab is in flight: compact the alignments with more than one CIGAR op (junctions_extractor.cc:379)
    {
        const uint4 o = *reinterpret_cast<const uint4*>(&sm.off[4 * t]);
        const uint32_t o4 = sm.off[4 * t + 4];
        const uint32_t r0 = 4 * t;
        uint32_t flags = 0;
        if (r0 + 0 < n_tile && o.y - o.x > 1u) flags |= 1u;
        if (r0 + 1 < n_tile && o.z - o.y > 1u) flags |= 2u;
        if (r0 + 2 < n_tile && o.w - o.z > 1u) flags |= 4u;
        if (r0 + 3 < n_tile && o4 - o.w > 1u) flags |= 8u;
        const uint32_t cnt = __popc(flags);
        uint32_t x = cnt;
#pragma unroll
        for (int dlt = 1; dlt < 32; dlt <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, dlt); if ((int)lane >= dlt) x += y; }
        uint32_t wbase = 0;
        if (lane == 31 && x) wbase = atomicAdd(&sm.n_work, x);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        uint32_t p = wbase + x - cnt;
        if (flags & 1u) sm.work[p++] = (uint16_t)(r0 + 0);
        if (flags & 2u) sm.work[p++] = (uint16_t)(r0 + 1);
        if (flags & 4u) sm.work[p++] = (uint16_t)(r0 + 2);
        if (flags & 8u) sm.work[p++] = (uint16_t)(r0 + 3);
    }
#pragma unroll
    for (int j = 0; j < S6_SV; ++j) {
        const uint32_t v = t + j * THREADS;
        if (4u * v < n_st) *reinterpret_cast<uint4*>(&sm.slab[4 * v]) = sv[j];
    }
    __syncthreads();
    const uint32_t n_work = (prm.debug & 1u) ? 0u : sm.n_work;
    if (n_work == 0) return;

    // ---- phase 3: walk, QC, combine, accumulate
    const FusedCtx cx{tb, spill, spill_cap, counters};
    const int32_t base_tid = (int32_t)sm.tid[0];
    const uint64_t ord0 = b.first_ordinal + base;
    uint32_t my_cands = 0;
    for (uint32_t w0 = 0; w0 < n_work; w0 += THREADS) {      // warp-uniform trip count
        const uint32_t i = w0 + t;
        uint32_t nc = 0, r = 0, sc = 0;
        int32_t tid = -1;
        WalkCand c0{0, 0, 0, 0, 0}, c1{0, 0, 0, 0, 0};
        if (i < n_work) {
            r = sm.work[i];
            tid = (int32_t)sm.tid[r];
            if (tid >= 0) {
                const uint32_t o0 = sm.off[r], n = sm.off[r + 1] - o0;
                sc = read_strand(sm.meta[r], prm.strandness);
                nc = (o0 - a0) + n <= n_st ? walk_collect<true>(sm.slab + (o0 - a0), n, sm.pos[r], c0, c1)
                                           : walk_collect<false>(b.cigar + o0, n, sm.pos[r], c0, c1);
                my_cands += nc;
                if (nc > 2u && !(prm.debug & 2u)) fused_walk_rest(cx, prm, ord0, b.cigar + o0, n, sm.pos[r], r, tid, sc);
            }
        }
        // junction_qc (:160-170) + add_junction (:174-235): warp-uniform passes over the (at most two) register
        // candidates of every lane
        for (uint32_t pass = 0; pass < 2u; ++pass) {
            bool has = nc > pass;
            if (!__any_sync(0xffffffffu, has)) break;
            const WalkCand c = pass ? c1 : c0;
            const uint32_t ilen = c.end - c.start;                                      // uint32, :161-162
            has = has && ilen >= prm.min_intron && ilen <= prm.max_intron;
            const uint32_t nts = ~(c.start - c.left), te = c.end + c.right;
            const uint32_t lr = (c.left >= prm.min_anchor ? 1u : 0u) | (c.right >= prm.min_anchor ? 2u : 0u);
            const uint32_t proxy = sc == '+' ? 0u : (sc == '-' ? 1u : 2u);               // :186-193
            const uint32_t nfirst_l = ~(r << 16 | c.k);
            const uint32_t last_l = proxy == 2u ? (r << 8 | sc) : 0u;
            if (prm.debug & 2u) { if (has && (nts ^ te ^ lr ^ nfirst_l ^ last_l) == 0x9e3779b9u) sm.sink = 1; continue; }
            if (has && (tid != base_tid || ilen >= (1u << 28))) {                      // not expressible in the block table's key
                fused_global_upsert(cx, ord0, c.start, c.end, tid, proxy, 1u, nts, te, lr, nfirst_l, last_l);
                has = false;
            }
            const unsigned long long key = (unsigned long long)c.start << 30 | (unsigned long long)ilen << 2 | proxy;
            // leader election: every distinct junction of the warp is reduced onto its first lane
            uint32_t todo = __ballot_sync(0xffffffffu, has);
            bool leader = false;
            uint32_t g_cnt = 0, g_nts = 0, g_te = 0, g_lr = 0, g_nf = 0, g_last = 0;
            while (todo) {
                const int src = __ffs(todo) - 1;
                const unsigned long long k = __shfl_sync(0xffffffffu, key, src);
                const bool in = has && key == k;
                const uint32_t grp = __ballot_sync(0xffffffffu, in);
                const uint32_t a1 = __reduce_max_sync(0xffffffffu, in ? nts : 0u), a2 = __reduce_max_sync(0xffffffffu, in ? te : 0u);
                const uint32_t a3 = __reduce_or_sync(0xffffffffu, in ? lr : 0u), a4 = __reduce_max_sync(0xffffffffu, in ? nfirst_l : 0u);
                const uint32_t a5 = __reduce_max_sync(0xffffffffu, in ? last_l : 0u);
                if ((int)lane == src) { leader = true; g_cnt = __popc(grp); g_nts = a1; g_te = a2; g_lr = a3; g_nf = a4; g_last = a5; }
                todo &= ~grp;
            }
            if (leader) {                                    // all leaders of the pass insert in parallel (distinct keys)
                uint32_t s = ((uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA77u) >> 25;   // 7 bits: S6_HS = 128
                bool done = false;
                for (int probe = 0; probe < S6_PROBES && !done; ++probe, s = (s + 1u) & (S6_HS - 1)) {
                    unsigned long long curk = sm.hkey[s];
                    if (curk == SKEY_EMPTY) curk = atomicCAS(&sm.hkey[s], SKEY_EMPTY, key);
                    if (curk == SKEY_EMPTY || curk == key) {
                        atomicAdd(&sm.hval[0][s], g_cnt);
                        atomicMax(&sm.hval[1][s], g_nts);
                        atomicMax(&sm.hval[2][s], g_te);
                        if (g_lr) atomicOr(&sm.hval[3][s], g_lr);
                        atomicMax(&sm.hval[4][s], g_nf);
                        if (g_last) atomicMax(&sm.hval[5][s], g_last);
                        done = true;
                    }
                }
                if (!done) fused_global_upsert(cx, ord0, c.start, c.end, tid, proxy, g_cnt, g_nts, g_te, g_lr, g_nf, g_last);
            }
        }
    }
    __syncthreads();
    // ---- phase 4: one global upsert per distinct junction of the tile
    if (!(prm.debug & 8u)) {
        for (uint32_t s = t; s < (uint32_t)S6_HS; s += THREADS) {
            const unsigned long long key = sm.hkey[s];
            if (key == SKEY_EMPTY) continue;
            const uint32_t start = (uint32_t)(key >> 30), ilen = (uint32_t)(key >> 2) & 0x0fffffffu;
            fused_global_upsert(cx, ord0, start, start + ilen, base_tid, (uint32_t)key & 3u, sm.hval[0][s], sm.hval[1][s], sm.hval[2][s],
                                sm.hval[3][s], sm.hval[4][s], sm.hval[5][s]);
        }
    }
    // statistics: N ops seen (before QC), one 64-bit RED per warp
    my_cands = __reduce_add_sync(0xffffffffu, my_cands);
    if (lane == 0 && my_cands) atomicAdd(reinterpret_cast<unsigned long long*>(counters + CTR_TOTAL_CAND64), (unsigned long long)my_cands);
}

// A/B probe (scan_debug bit 5): reads the five arrays of the batch once with plain 128-bit loads and nothing else —
// the practical read ceiling of the device for a batch of this size, next to which cigar_scan's time is judged.
__global__ void __launch_bounds__(256)
stream_probe_kernel(BatchView b, uint32_t* __restrict__ counters) {
    const uint4* cols[5] = {reinterpret_cast<const uint4*>(b.tid), reinterpret_cast<const uint4*>(b.pos),
                            reinterpret_cast<const uint4*>(b.meta), reinterpret_cast<const uint4*>(b.cig_off),
                            reinterpret_cast<const uint4*>(b.cigar)};
    const size_t nvec[5] = {b.n_reads / 4u, b.n_reads / 4u, b.n_reads / 4u, b.n_reads / 4u, b.n_ops / 4u};
    uint32_t acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        size_t i = t0;
        for (; i + 3 * stride < nvec[c]; i += 4 * stride) {
            const uint4 x0 = ldg_stream_u4(cols[c] + i), x1 = ldg_stream_u4(cols[c] + i + stride);
            const uint4 x2 = ldg_stream_u4(cols[c] + i + 2 * stride), x3 = ldg_stream_u4(cols[c] + i + 3 * stride);
            acc ^= x0.x ^ x0.y ^ x0.z ^ x0.w ^ x1.x ^ x1.y ^ x1.z ^ x1.w ^ x2.x ^ x2.y ^ x2.z ^ x2.w ^ x3.x ^ x3.y ^ x3.z ^ x3.w;
        }
        for (; i < nvec[c]; i += stride) { const uint4 x = ldg_stream_u4(cols[c] + i); acc ^= x.x ^ x.y ^ x.z ^ x.w; }
    }
    if (acc == 0x9e3779b9u) atomicAdd(&counters[CTR_NOUT], 1u);
}

// A/B probe (scan_debug bit 6): the SAME access pattern as cigar_scan (512-alignment tiles: 2 KB of each of the four
// columns + the tile's CIGAR slab, tiles strided over a persistent grid) but with plain register loads and no shared
// memory: separates "the pattern" from "the cp.async staging".
template <bool STAGE>
__global__ void __launch_bounds__(128)
tile_probe_kernel(BatchView b, uint32_t* __restrict__ counters) {
    __shared__ uint4 stage[STAGE ? 680 : 1];
    const uint32_t t = threadIdx.x, n_tiles = b.n_reads / 512u;
    uint32_t acc = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t base = tile * 512u;
        const uint4 o = ldg_stream_u4(reinterpret_cast<const uint4*>(b.cig_off + base) + t);
        const uint4 p = ldg_stream_u4(reinterpret_cast<const uint4*>(b.pos + base) + t);
        const uint4 m = ldg_stream_u4(reinterpret_cast<const uint4*>(b.meta + base) + t);
        const uint4 d = ldg_stream_u4(reinterpret_cast<const uint4*>(b.tid + base) + t);
        const uint32_t lo = __shfl_sync(0xffffffffu, o.x, 0);           // warp 0's lane 0 holds cig_off[base]; good enough for a probe
        const uint32_t a0 = (lo & ~3u) + 4u * t;
        uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
        if (a0 + 4u <= b.n_ops) c0 = ldg_stream_u4(reinterpret_cast<const uint4*>(b.cigar + a0));
        if (t < 40u && a0 + 516u <= b.n_ops) c1 = ldg_stream_u4(reinterpret_cast<const uint4*>(b.cigar + a0 + 512u));
        if (STAGE) {                                          // register-staged copy into shared memory + block barrier per tile
            __syncthreads();
            stage[t] = o; stage[128 + t] = p; stage[256 + t] = m; stage[384 + t] = d; stage[512 + t] = c0;
            if (t < 40u) stage[640 + t] = c1;
            __syncthreads();
            acc ^= stage[(t * 7u) % 680u].y;
        } else {
            acc ^= o.x ^ o.w ^ p.x ^ p.w ^ m.x ^ m.w ^ d.x ^ d.w ^ c0.x ^ c0.w ^ c1.x ^ c1.w;
        }
    }
    if (acc == 0x9e3779b9u) atomicAdd(&counters[CTR_NOUT], 1u);
}

// A/B probe (scan_debug bit 18 with bit 6; hint in bits 16-17): the tile pattern through cp.async (LDGSTS) into a 2-deep shared-memory
// ring, with an optional L2 prefetch-size hint on the copies.  HINT: 0 none, 1 L2::128B, 2 L2::256B.
template <int HINT>
__device__ __forceinline__ void cp_async16_hint(void* dst, const void* src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    if (HINT == 2) asm volatile("cp.async.cg.shared.global.L2::256B [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else if (HINT == 1) asm volatile("cp.async.cg.shared.global.L2::128B [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
    else asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
template <int HINT>
__global__ void __launch_bounds__(128)
tile_probe_async_kernel(BatchView b, uint32_t* __restrict__ counters) {
    __shared__ uint4 ring[2][768];
    const uint32_t t = threadIdx.x, n_tiles = b.n_reads / 512u;
    const uint32_t words_per_tile = (b.n_ops / n_tiles) & ~3u;          // stand-in for the slab address: no dependent load in a probe
    uint32_t acc = 0, it = 0;
    auto issue = [&](uint32_t tile, uint4* buf) {
        const uint32_t base = tile * 512u;
        cp_async16_hint<HINT>(buf + t, reinterpret_cast<const uint4*>(b.cig_off + base) + t);
        cp_async16_hint<HINT>(buf + 128 + t, reinterpret_cast<const uint4*>(b.pos + base) + t);
        cp_async16_hint<HINT>(buf + 256 + t, reinterpret_cast<const uint4*>(b.meta + base) + t);
        cp_async16_hint<HINT>(buf + 384 + t, reinterpret_cast<const uint4*>(b.tid + base) + t);
        const uint32_t a0 = tile * words_per_tile + 4u * t;
        if (a0 + 4u <= b.n_ops) cp_async16_hint<HINT>(buf + 512 + t, b.cigar + a0);
        if (t < 40u && a0 + 516u <= b.n_ops) cp_async16_hint<HINT>(buf + 640 + t, b.cigar + a0 + 512u);
    };
    if (blockIdx.x < n_tiles) issue(blockIdx.x, ring[0]);
    cp_async_commit();
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        if (tile + gridDim.x < n_tiles) issue(tile + gridDim.x, ring[(it + 1) & 1]);
        cp_async_commit();
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        acc ^= ring[it & 1][(t * 7u) % 680u].y;
        __syncthreads();
    }
    if (acc == 0x9e3779b9u) atomicAdd(&counters[CTR_NOUT], 1u);
}

template <int THREADS, int MIN_BLOCKS>
static void launch_fused_cfg(const BatchView& b, const ScanParams& p, const TableRef& tb, Slot* spill, uint32_t spill_cap,
                             uint32_t* d_counters, cudaStream_t stream) {
    static bool once = false;
    if (!once) {
        once = true;
        cudaFuncSetAttribute(cigar_scan_fused_kernel<THREADS, MIN_BLOCKS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (getenv("RTJX_TRACE")) {
            int nb = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, cigar_scan_fused_kernel<THREADS, MIN_BLOCKS>, THREADS, 0);
            fprintf(stderr, "[rtjx] cigar_scan_fused<%d,%d>: %zu B shared memory, %d blocks/SM, %d SMs\n", THREADS, MIN_BLOCKS,
                    sizeof(S6Smem<THREADS>), nb, num_sms());
        }
    }
    const uint32_t tiles = (b.n_reads + THREADS * 4 - 1) / (THREADS * 4);
    cigar_scan_fused_kernel<THREADS, MIN_BLOCKS><<<tiles, THREADS, 0, stream>>>(b, p, tb, spill, spill_cap, d_counters);
}

bool launch_cigar_scan_fused(const BatchView& b, const ScanParams& p, const TableRef& tb, Slot* spill, uint32_t spill_cap,
                             uint32_t* d_counters, cudaStream_t stream) {
    if (b.n_reads == 0) return true;
    const uintptr_t align = reinterpret_cast<uintptr_t>(b.tid) | reinterpret_cast<uintptr_t>(b.pos) |
                            reinterpret_cast<uintptr_t>(b.meta) | reinterpret_cast<uintptr_t>(b.cig_off) |
                            reinterpret_cast<uintptr_t>(b.cigar);
    if (align & 15u) return false;
    if (p.debug & 32u) { stream_probe_kernel<<<num_sms() * 8, 256, 0, stream>>>(b, d_counters); return true; }
    if (p.debug & 64u) {
        const uint32_t per_sm = ((p.debug >> 8) & 0xffu) ? ((p.debug >> 8) & 0xffu) : 16u;
        if (p.debug & (1u << 18)) {
            const uint32_t hint = (p.debug >> 16) & 3u;
            if (hint == 2) tile_probe_async_kernel<2><<<num_sms() * per_sm, 128, 0, stream>>>(b, d_counters);
            else if (hint == 1) tile_probe_async_kernel<1><<<num_sms() * per_sm, 128, 0, stream>>>(b, d_counters);
            else tile_probe_async_kernel<0><<<num_sms() * per_sm, 128, 0, stream>>>(b, d_counters);
        } else if (p.debug & 128u) tile_probe_kernel<true><<<num_sms() * per_sm, 128, 0, stream>>>(b, d_counters);
        else tile_probe_kernel<false><<<num_sms() * per_sm, 128, 0, stream>>>(b, d_counters);
        return true;
    }
    switch (p.cfg) {            // A/B configurations; 0 is the production one
    case 1: launch_fused_cfg<128, 10>(b, p, tb, spill, spill_cap, d_counters, stream); break;
    case 2: launch_fused_cfg<64, 16>(b, p, tb, spill, spill_cap, d_counters, stream); break;
    case 3: launch_fused_cfg<256, 4>(b, p, tb, spill, spill_cap, d_counters, stream); break;
    case 4: launch_fused_cfg<128, 6>(b, p, tb, spill, spill_cap, d_counters, stream); break;
    default: launch_fused_cfg<128, 8>(b, p, tb, spill, spill_cap, d_counters, stream); break;
    }
    return true;
}

// Re-inserts every occupied slot of `src` (an old table, or the spill list) into the table.
__global__ void __launch_bounds__(256)
table_rehash_kernel(const Slot* __restrict__ src, uint32_t n_src, TableRef tb, uint32_t* __restrict__ counters) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
        Slot s = src[i];
        if (s.khi == 0ull) continue;
        if (!table_upsert(tb, K128{s.klo, s.khi}, s.count, s.nts, s.te, s.lr, s.nfirst, s.last, counters))
            atomicExch(&counters[CTR_CAND_OVERFLOW], 2u);
    }
}

void launch_table_rehash(const Slot* old_table, uint32_t old_slots, const TableRef& tb, uint32_t* d_counters,
                         cudaStream_t stream) {
    if (old_slots == 0) return;
    uint32_t grid = min((old_slots + 255u) / 256u, (uint32_t)(8 * num_sms()));
    table_rehash_kernel<<<grid, 256, 0, stream>>>(old_table, old_slots, tb, d_counters);
}

// `-b` mode: pair table -> junction table (see jx_device.cuh).  Only the occupied slots are visited (slot_list).
__global__ void __launch_bounds__(256)
table_fold_kernel(TableRef src, uint32_t n, TableRef dst, uint32_t* __restrict__ dst_counters) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const Slot s = src.slots[src.slot_list[i]];
        if (s.khi == 0ull) continue;
        if (!table_upsert(dst, K128{s.klo, s.khi & ((1ull << 34) - 1ull)}, s.count, s.nts, s.te, s.lr, s.nfirst, s.last, dst_counters))
            atomicExch(&dst_counters[CTR_CAND_OVERFLOW], 2u);
    }
}

void launch_table_fold(const TableRef& src, uint32_t n, const TableRef& dst, uint32_t* dst_counters, cudaStream_t stream) {
    if (n == 0) return;
    uint32_t grid = min((n + 255u) / 256u, (uint32_t)(8 * num_sms()));
    table_fold_kernel<<<grid, 256, 0, stream>>>(src, n, dst, dst_counters);
}

// Zeroes the occupied slots (rtjx_clear); 3 x 16 bytes per slot.
__global__ void __launch_bounds__(256)
table_clear_kernel(TableRef tb, const uint32_t* __restrict__ d_n_unique) {
    const uint32_t n = min(*d_n_unique, tb.list_cap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint4* p = reinterpret_cast<uint4*>(tb.slots + tb.slot_list[i]);
        p[0] = make_uint4(0, 0, 0, 0); p[1] = make_uint4(0, 0, 0, 0); p[2] = make_uint4(0, 0, 0, 0);
    }
}

void launch_table_clear(const TableRef& tb, const uint32_t* d_n_unique, uint32_t n_bound, cudaStream_t stream) {
    if (n_bound == 0) return;
    uint32_t grid = min((n_bound + 255u) / 256u, (uint32_t)(8 * num_sms()));
    table_clear_kernel<<<grid, 256, 0, stream>>>(tb, d_n_unique);
}

// ------------------------------------------------------------------------------------------------
// finalize: compaction, first-seen ranking, sort
// ------------------------------------------------------------------------------------------------
// slot_list[i] -> OutJunction[i]; no scan of the (mostly empty) table.
__global__ void __launch_bounds__(256)
table_compact_kernel(TableRef tb, uint32_t n, OutJunction* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* p = reinterpret_cast<const uint4*>(tb.slots + tb.slot_list[i]);
    const uint4 k = p[0], v = p[1], w = p[2];
    const unsigned long long khi = (unsigned long long)k.w << 32 | k.z;
    const unsigned long long last = (unsigned long long)w.w << 32 | w.z;
    const uint32_t proxy = (uint32_t)khi & 3u;
    OutJunction j;
    j.tid = (int32_t)(uint32_t)(khi >> 2) - 1;          // (uint32_t) drops the variant-region bits 34..
    j.start = k.y; j.end = k.x;                      // klo = start << 32 | end
    j.ts = ~v.y; j.te = v.z; j.count = v.x; j.name_index = 0;
    j.strand = proxy == 0u ? '+' : (proxy == 1u ? '-' : (uint8_t)(last & 0xffu));
    j.left_ok = v.w & 1u; j.right_ok = (v.w >> 1) & 1u; j.pad = 0;
    j.first_ord = ~((unsigned long long)w.y << 32 | w.x);
    out[i] = j;
}

void launch_table_compact(const TableRef& tb, uint32_t n, OutJunction* out, cudaStream_t stream) {
    if (n == 0) return;
    table_compact_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(tb, n, out);
}

struct ByFirstOrd {
    __device__ __forceinline__ bool operator()(const OutJunction& a, const OutJunction& b) const { return a.first_ord < b.first_ord; }
};
// compare_junctions (junctions_extractor.h:117-140) with the contig string order precomputed as a rank
struct ByBedOrder {
    const uint32_t* contig_rank; uint32_t n_contigs;
    __device__ __forceinline__ uint32_t cr(int32_t tid) const {
        return (uint32_t)tid < n_contigs ? contig_rank[tid] : 0x40000000u + (uint32_t)tid;
    }
    __device__ __forceinline__ bool operator()(const OutJunction& a, const OutJunction& b) const {
        const uint32_t ca = cr(a.tid), cb = cr(b.tid);
        if (ca != cb) return ca < cb;
        if (a.ts != b.ts) return a.ts < b.ts;
        if (a.te != b.te) return a.te < b.te;
        return a.name_index < b.name_index;
    }
};
__global__ void fin_assign_names(OutJunction* __restrict__ e, uint32_t n) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) e[r].name_index = r + 1u;            // rank of first appearance (junctions_extractor.cc:152-157)
}

size_t finalize_sort_workspace_bytes(uint32_t n) {
    size_t a = 0, b = 0;
    cub::DeviceMergeSort::SortKeys(nullptr, a, (OutJunction*)nullptr, (int)n, ByFirstOrd());
    cub::DeviceMergeSort::SortKeys(nullptr, b, (OutJunction*)nullptr, (int)n, ByBedOrder{nullptr, 0});
    return (a > b ? a : b) + 256;
}

// entries[0..n): ranked by first_ord (name_index), then sorted in place by (contig string, ts, te, name).
void launch_finalize_sort(OutJunction* entries, uint32_t n, const uint32_t* contig_rank, uint32_t n_contigs,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n == 0) return;
    cub::DeviceMergeSort::SortKeys(workspace, workspace_bytes, entries, (int)n, ByFirstOrd(), stream);
    fin_assign_names<<<(n + 255u) / 256u, 256, 0, stream>>>(entries, n);
    cub::DeviceMergeSort::SortKeys(workspace, workspace_bytes, entries, (int)n, ByBedOrder{contig_rank, n_contigs}, stream);
}

// ------------------------------------------------------------------------------------------------
// finalize, batched variant-region mode
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
table_compact_regions_kernel(TableRef tb, uint32_t n, OutJunctionR* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* p = reinterpret_cast<const uint4*>(tb.slots + tb.slot_list[i]);
    const uint4 k = p[0], v = p[1], w = p[2];
    const unsigned long long khi = (unsigned long long)k.w << 32 | k.z;
    const unsigned long long last = (unsigned long long)w.w << 32 | w.z;
    const uint32_t proxy = (uint32_t)khi & 3u;
    OutJunctionR r;
    r.j.tid = (int32_t)((uint32_t)(khi >> 2)) - 1;
    r.j.start = k.y; r.j.end = k.x;
    r.j.ts = ~v.y; r.j.te = v.z; r.j.count = v.x; r.j.name_index = 0;
    r.j.strand = proxy == 0u ? '+' : (proxy == 1u ? '-' : (uint8_t)(last & 0xffu));
    r.j.left_ok = v.w & 1u; r.j.right_ok = (v.w >> 1) & 1u; r.j.pad = 0;
    r.j.first_ord = ~((unsigned long long)w.y << 32 | w.x);
    r.region = (uint32_t)(khi >> 34); r.pad = 0;          // region index + 1
    out[i] = r;
}
void launch_table_compact_regions(const TableRef& tb, uint32_t n, OutJunctionR* out, cudaStream_t stream) {
    if (n == 0) return;
    table_compact_regions_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(tb, n, out);
}
struct ByRegionFirstOrd {
    __device__ __forceinline__ bool operator()(const OutJunctionR& a, const OutJunctionR& b) const {
        return a.region != b.region ? a.region < b.region : a.j.first_ord < b.j.first_ord;
    }
};
struct ByRegionBedOrder {
    ByBedOrder inner;
    __device__ __forceinline__ bool operator()(const OutJunctionR& a, const OutJunctionR& b) const {
        return a.region != b.region ? a.region < b.region : inner(a.j, b.j);
    }
};
// every region is its own extractor: JUNC numbering restarts at 1 (junctions_extractor.cc:152-157 on a fresh object)
__global__ void fin_assign_names_regions(OutJunctionR* __restrict__ e, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t rg = e[i].region;
    uint32_t lo = 0, hi = i;                              // first index of this region in the (region, first_ord)-sorted array
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (e[mid].region < rg) lo = mid + 1; else hi = mid; }
    e[i].j.name_index = i - lo + 1u;
}
// `-b` mode: runs of one junction key, barcodes in first-seen order (the order the reference inserted them, :203-215)
struct ByJunctionFirstOrd {
    __device__ __forceinline__ bool operator()(const OutJunctionR& a, const OutJunctionR& b) const {
        if (a.j.tid != b.j.tid) return a.j.tid < b.j.tid;
        if (a.j.start != b.j.start) return a.j.start < b.j.start;
        if (a.j.end != b.j.end) return a.j.end < b.j.end;
        const uint32_t pa = a.j.strand == '+' ? 0u : (a.j.strand == '-' ? 1u : 2u), pb = b.j.strand == '+' ? 0u : (b.j.strand == '-' ? 1u : 2u);
        if (pa != pb) return pa < pb;
        return a.j.first_ord < b.j.first_ord;
    }
};
void launch_sort_barcode_pairs(OutJunctionR* entries, uint32_t n, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n == 0) return;
    cub::DeviceMergeSort::SortKeys(workspace, workspace_bytes, entries, (int)n, ByJunctionFirstOrd(), stream);
}
size_t finalize_sort_regions_workspace_bytes(uint32_t n) {
    size_t a = 0, b = 0, c = 0;
    cub::DeviceMergeSort::SortKeys(nullptr, c, (OutJunctionR*)nullptr, (int)n, ByJunctionFirstOrd());
    cub::DeviceMergeSort::SortKeys(nullptr, a, (OutJunctionR*)nullptr, (int)n, ByRegionFirstOrd());
    cub::DeviceMergeSort::SortKeys(nullptr, b, (OutJunctionR*)nullptr, (int)n, ByRegionBedOrder{ByBedOrder{nullptr, 0}});
    if (c > a) a = c;
    return (a > b ? a : b) + 256;
}
void launch_finalize_sort_regions(OutJunctionR* entries, uint32_t n, const uint32_t* contig_rank, uint32_t n_contigs,
                                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n == 0) return;
    cub::DeviceMergeSort::SortKeys(workspace, workspace_bytes, entries, (int)n, ByRegionFirstOrd(), stream);
    fin_assign_names_regions<<<(n + 255u) / 256u, 256, 0, stream>>>(entries, n);
    cub::DeviceMergeSort::SortKeys(workspace, workspace_bytes, entries, (int)n, ByRegionBedOrder{ByBedOrder{contig_rank, n_contigs}}, stream);
}

}  // namespace rtjx
