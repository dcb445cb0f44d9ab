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
#include <cub/device/device_radix_sort.cuh>

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

// ------------------------------------------------------------------------------------------------
// cigar_scan, persistent TMA-pipelined version (the one normally launched)
// ------------------------------------------------------------------------------------------------
// One CTA per half-SM loops over tiles of 1024 alignments.  The four metadata columns of a tile
// (pos, meta, tid, cig_off: 16 KB) and its contiguous CIGAR slab are brought into a 3-stage shared
// memory ring by 1-D bulk async copies (cp.async.bulk / UBLKCP) issued by one thread and tracked
// by mbarriers, so the loads of tiles j+1 and j+2 are in flight while tile j is processed: the SM
// always has tens of KB outstanding to HBM without spending issue slots on LDGs.  Only ~10 % of
// alignments have more than one CIGAR op; a ballot/prefix-sum pass compacts their indices into a
// shared work list so that the op walk runs on dense warps instead of diverging in every warp.
constexpr int S2_THREADS = 256;
constexpr int S2_TILE    = 1024;
constexpr int S2_STAGES  = 3;
constexpr int S2_SLAB    = 3072;                 // CIGAR words per stage (12 KB)
constexpr int S2_OUT     = 512;                  // staged candidates per tile (16 KB)

struct alignas(16) S2Stage {
    uint32_t pos[S2_TILE];
    uint32_t meta[S2_TILE];
    uint32_t tid[S2_TILE];
    uint32_t off[S2_TILE + 4];
    uint32_t slab[S2_SLAB];
};
struct alignas(16) S2Smem {
    S2Stage st[S2_STAGES];
    uint4 out[S2_OUT * 2];
    unsigned long long meta_full[S2_STAGES];
    unsigned long long slab_full[S2_STAGES];
    uint32_t slab_a0[S2_STAGES];                 // first staged word index (16-byte aligned)
    uint32_t slab_direct[S2_STAGES];             // 1: slab not staged, walk reads global memory
    uint32_t n_work[2], n_out[2];
    uint32_t flush_base;
    uint16_t work[S2_TILE];
};

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

struct S2Emit {
    S2Smem& sm; uint32_t buf; Cand* __restrict__ out; uint32_t cap; uint32_t* counters;
    __device__ __forceinline__ void operator()(uint32_t start, uint32_t end, uint32_t left, uint32_t right,
                                               uint64_t ord, int32_t tid, uint32_t strand) const {
        uint4 a = make_uint4(start, end, start - left, end + right);
        uint4 b = make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)tid, strand);
        uint32_t i = atomicAdd(&sm.n_out[buf], 1u);
        if (i < S2_OUT) {
            sm.out[2 * i] = a; sm.out[2 * i + 1] = b;
        } else {
            uint32_t g = atomicAdd(&counters[CTR_NCAND], 1u);
            if (g < cap) { uint4* o = reinterpret_cast<uint4*>(out + g); o[0] = a; o[1] = b; }
            else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
        }
    }
};

// closed-form walk, same arithmetic as scan_walk above
template <bool FROM_SMEM, class Emit>
__device__ __forceinline__ void walk_ops(const uint32_t* __restrict__ ops, uint32_t n, uint32_t pos, int32_t tid,
                                         uint32_t strand, uint64_t read_ord, const Emit& emit) {
    const uint32_t ANC = (1u << 0) | (1u << 7);
    const uint32_t BRK = (1u << 3) | (1u << 2) | (1u << 8) | (1u << 1) | (1u << 4);
    const uint32_t REFC = ANC | (1u << 2) | (1u << 8) | (1u << 3);
    uint32_t cur = pos, run = 0;
    bool pending = false;
    uint32_t p_start = 0, p_end = 0, p_left = 0, p_k = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint32_t w = FROM_SMEM ? ops[i] : __ldg(ops + i);
        const uint32_t op = w & 0xfu, len = w >> 4, bit = 1u << op;
        if (bit & BRK) {
            if (pending) { emit(p_start, p_end, p_left, run, read_ord << 16 | (p_k > 0xffffu ? 0xffffu : p_k), tid, strand); pending = false; }
            if (op == 3u) { pending = true; p_start = cur; p_end = cur + len; p_left = run; p_k = i; }
            run = 0;
        } else if (bit & ANC) {
            run += len;
        }
        if (bit & REFC) cur += len;
    }
    if (pending) emit(p_start, p_end, p_left, run, read_ord << 16 | (p_k > 0xffffu ? 0xffffu : p_k), tid, strand);
}

__global__ void __launch_bounds__(S2_THREADS, 2)
cigar_scan_tma_kernel(BatchView b, ScanParams prm, Cand* __restrict__ out, uint32_t cap, uint32_t* __restrict__ counters) {
    extern __shared__ __align__(128) unsigned char s2_raw[];
    S2Smem& sm = *reinterpret_cast<S2Smem*>(s2_raw);
    const uint32_t t = threadIdx.x, lane = t & 31u;
    const uint32_t n_tiles = (b.n_reads + S2_TILE - 1) / S2_TILE;
    const uint32_t n_ops_vec_end = b.n_ops & ~3u;          // bulk copies must not run past the array

    if (t == 0) {
        for (int s = 0; s < S2_STAGES; ++s) { mbar_init(&sm.meta_full[s], 1); mbar_init(&sm.slab_full[s], 1); }
        sm.n_work[0] = sm.n_work[1] = 0; sm.n_out[0] = sm.n_out[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // metadata of `tile` -> stage s.  Full tiles whose cig_off window (1028 entries) lies inside the
    // array use four bulk copies from one thread; the (at most two) ragged tiles at the end of the
    // batch are loaded cooperatively.  Block-uniform; contains a barrier only on the ragged path.
    auto issue_meta = [&](uint32_t tile, int s) {
        const uint32_t base = tile * S2_TILE;
        S2Stage& st = sm.st[s];
        if (base + S2_TILE + 3 <= b.n_reads) {
            if (t == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive_expect_tx(&sm.meta_full[s], 3 * S2_TILE * 4 + (S2_TILE + 4) * 4);
                bulk_g2s(st.off, b.cig_off + base, (S2_TILE + 4) * 4, &sm.meta_full[s]);
                bulk_g2s(st.pos, b.pos + base, S2_TILE * 4, &sm.meta_full[s]);
                bulk_g2s(st.meta, b.meta + base, S2_TILE * 4, &sm.meta_full[s]);
                bulk_g2s(st.tid, b.tid + base, S2_TILE * 4, &sm.meta_full[s]);
            }
        } else {
            const uint32_t n_tile = min((uint32_t)S2_TILE, b.n_reads - base);
            for (uint32_t r = t; r < n_tile; r += S2_THREADS) {
                st.pos[r] = (uint32_t)b.pos[base + r]; st.meta[r] = b.meta[base + r]; st.tid[r] = (uint32_t)b.tid[base + r];
            }
            for (uint32_t r = t; r <= n_tile; r += S2_THREADS) st.off[r] = b.cig_off[base + r];
            __syncthreads();
            if (t == 0) mbar_arrive(&sm.meta_full[s]);
        }
    };
    // CIGAR slab of the tile whose metadata is (or will be) in stage s.  Thread 0 only.
    auto issue_slab = [&](uint32_t tile, int s, uint32_t parity) {
        mbar_wait(&sm.meta_full[s], parity);
        S2Stage& st = sm.st[s];
        const uint32_t n_tile = min((uint32_t)S2_TILE, b.n_reads - tile * S2_TILE);
        const uint32_t lo = st.off[0], hi = st.off[n_tile];
        const uint32_t a0 = lo & ~3u, end4 = (hi + 3u) & ~3u;
        sm.slab_a0[s] = a0;
        if (hi <= lo) {                                       // no ops at all
            sm.slab_direct[s] = 0; mbar_arrive(&sm.slab_full[s]);
        } else if (end4 - a0 > (uint32_t)S2_SLAB || end4 > n_ops_vec_end) {
            sm.slab_direct[s] = 1; mbar_arrive(&sm.slab_full[s]);
        } else {
            sm.slab_direct[s] = 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_expect_tx(&sm.slab_full[s], (end4 - a0) * 4);
            bulk_g2s(st.slab, b.cigar + a0, (end4 - a0) * 4, &sm.slab_full[s]);
        }
    };

    // prologue: metadata of the first S2_STAGES tiles of this CTA, slab of the first
    for (int s = 0; s < S2_STAGES; ++s) {
        const uint32_t tile = blockIdx.x + (uint32_t)s * gridDim.x;
        if (tile < n_tiles) issue_meta(tile, s);
    }
    if (t == 0 && blockIdx.x < n_tiles) issue_slab(blockIdx.x, 0, 0);

    uint32_t k = 0;
    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++k) {
        const int s = (int)(k % S2_STAGES);
        const uint32_t parity = (k / S2_STAGES) & 1u;
        const uint32_t cb = k & 1u;                          // counter buffer of this iteration
        const uint32_t next = tile + gridDim.x;
        if (t == 0 && next < n_tiles) issue_slab(next, (int)((k + 1) % S2_STAGES), ((k + 1) / S2_STAGES) & 1u);
        mbar_wait(&sm.meta_full[s], parity);
        mbar_wait(&sm.slab_full[s], parity);
        S2Stage& st = sm.st[s];
        const uint32_t base = tile * S2_TILE;
        const uint32_t n_tile = min((uint32_t)S2_TILE, b.n_reads - base);

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
            if (flags & 1u) sm.work[p++] = (uint16_t)(r0 + 0);
            if (flags & 2u) sm.work[p++] = (uint16_t)(r0 + 1);
            if (flags & 4u) sm.work[p++] = (uint16_t)(r0 + 2);
            if (flags & 8u) sm.work[p++] = (uint16_t)(r0 + 3);
        }
        __syncthreads();
        if (t == 0) { sm.n_work[cb ^ 1u] = 0; sm.n_out[cb ^ 1u] = 0; }

        // ---- phase B: walk the compacted alignments, one per thread
        {
            const uint32_t n_work = sm.n_work[cb];
            const uint32_t a0 = sm.slab_a0[s];
            const bool direct = sm.slab_direct[s] != 0;
            const S2Emit emit{sm, cb, out, cap, counters};
            for (uint32_t w = t; w < n_work; w += S2_THREADS) {
                const uint32_t r = sm.work[w];
                const int32_t tid = (int32_t)st.tid[r];
                if (tid < 0) continue;
                const uint32_t o0 = st.off[r], n = st.off[r + 1] - o0;
                const uint32_t strand = read_strand(st.meta[r], prm.strandness);
                const uint64_t read_ord = b.first_ordinal + base + r;
                if (!direct) walk_ops<true>(st.slab + (o0 - a0), n, st.pos[r], tid, strand, read_ord, emit);
                else walk_ops<false>(b.cigar + o0, n, st.pos[r], tid, strand, read_ord, emit);
            }
        }
        __syncthreads();

        // ---- stage s is drained: refill it with the tile S2_STAGES iterations ahead
        {
            const uint32_t refill = tile + (uint32_t)S2_STAGES * gridDim.x;
            if (refill < n_tiles) issue_meta(refill, s);
        }
        // ---- flush the staged candidates with one reservation per tile
        const uint32_t n_st = min(sm.n_out[cb], (uint32_t)S2_OUT);
        if (n_st) {
            if (t == 0) sm.flush_base = atomicAdd(&counters[CTR_NCAND], n_st);
            __syncthreads();
            const uint32_t fb = sm.flush_base;
            uint4* o = reinterpret_cast<uint4*>(out);
            for (uint32_t v = t; v < 2 * n_st; v += S2_THREADS) {
                if (fb + (v >> 1) < cap) o[2ull * fb + v] = sm.out[v];
                else atomicExch(&counters[CTR_CAND_OVERFLOW], 1u);
            }
        }
    }
}

void launch_cigar_scan(const BatchView& b, const ScanParams& p, Cand* cands, uint32_t cand_cap,
                       uint32_t* d_counters, cudaStream_t stream) {
    if (b.n_reads == 0) return;
    const uintptr_t align = reinterpret_cast<uintptr_t>(b.tid) | reinterpret_cast<uintptr_t>(b.pos) |
                            reinterpret_cast<uintptr_t>(b.meta) | reinterpret_cast<uintptr_t>(b.cig_off) |
                            reinterpret_cast<uintptr_t>(b.cigar);
    if ((align & 15u) == 0) {            // bulk async copies need 16-byte aligned columns (cudaMalloc gives 256)
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(cigar_scan_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S2Smem));
            attr_set = true;
        }
        const uint32_t tiles = (b.n_reads + S2_TILE - 1) / S2_TILE;
        const uint32_t grid = min(tiles, (uint32_t)(2 * num_sms()));
        cigar_scan_tma_kernel<<<grid, S2_THREADS, sizeof(S2Smem), stream>>>(b, p, cands, cand_cap, d_counters);
        return;
    }
    uint32_t grid = (b.n_reads + SCAN_TILE - 1) / SCAN_TILE;
    cigar_scan_kernel<<<grid, SCAN_THREADS, 0, stream>>>(b, p, cands, cand_cap, d_counters);
}

// ------------------------------------------------------------------------------------------------
// device-wide junction table
// ------------------------------------------------------------------------------------------------
// Upsert of an (already aggregated) partial reduction into the global table.  Returns false if the
// table has no free slot on the probe path (caller spills).
__device__ __forceinline__ bool table_upsert(Slot* __restrict__ table, uint32_t mask, K128 key, uint32_t count,
                                             uint32_t nts, uint32_t te, uint32_t lr, unsigned long long nfirst,
                                             unsigned long long last, uint32_t* counters) {
    uint32_t s = mix_key(key.lo, key.hi) & mask;
    for (uint32_t probe = 0; probe <= mask; ++probe, s = (s + 1) & mask) {
        Slot* sl = table + s;
        K128 cur = ld128_relaxed(sl);
        if (cur.lo == 0ull && cur.hi == 0ull) {
            cur = cas128(sl, K128{0ull, 0ull}, key);
            if (cur.lo == 0ull && cur.hi == 0ull) {
                atomicAdd(&counters[CTR_NUNIQUE], 1u);
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

struct MergeSmem {
    unsigned long long key[MERGE_SLOTS];
    unsigned long long nfirst[MERGE_SLOTS];
    unsigned long long last[MERGE_SLOTS];
    uint32_t count[MERGE_SLOTS], nts[MERGE_SLOTS], te[MERGE_SLOTS], lr[MERGE_SLOTS];
    int32_t base_tid;
};

__global__ void __launch_bounds__(MERGE_THREADS, 2)
junction_merge_kernel(const Cand* __restrict__ cands, const uint32_t* __restrict__ d_n_cand, uint32_t n_bound,
                      ScanParams prm, Slot* __restrict__ table, uint32_t mask, Slot* __restrict__ spill,
                      uint32_t spill_cap, uint32_t* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MergeSmem& sm = *reinterpret_cast<MergeSmem*>(smem_raw);
    const uint32_t t = threadIdx.x;
    uint32_t n = d_n_cand ? min(*d_n_cand, n_bound) : n_bound;
    const uint32_t n_tiles = (n + MERGE_TILE - 1) / MERGE_TILE;
    if (blockIdx.x == 0 && t == 0)
        atomicAdd(reinterpret_cast<unsigned long long*>(counters + CTR_TOTAL_CAND64), (unsigned long long)n);

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t tb = tile * MERGE_TILE;
        for (uint32_t s = t; s < MERGE_SLOTS; s += MERGE_THREADS) {
            sm.key[s] = SKEY_EMPTY; sm.nfirst[s] = 0ull; sm.last[s] = 0ull;
            sm.count[s] = 0u; sm.nts[s] = 0u; sm.te[s] = 0u; sm.lr[s] = 0u;
        }
        if (t == 0) sm.base_tid = cands[tb].tid;
        __syncthreads();
        const int32_t base_tid = sm.base_tid;

        // all loads of the tile first (two 128-bit loads per candidate, coalesced across the warp)
        uint4 ca[MERGE_CPT], cb[MERGE_CPT];
#pragma unroll
        for (int j = 0; j < MERGE_CPT; ++j) {
            uint32_t i = tb + t + j * MERGE_THREADS;
            if (i < n) {
                const uint4* p = reinterpret_cast<const uint4*>(cands + i);
                ca[j] = ldg_stream_u4(p);
                cb[j] = ldg_stream_u4(p + 1);
            } else {
                ca[j] = make_uint4(0, 0, 0, 0);
                cb[j] = make_uint4(0, 0, 0xffffffffu, 0);      // tid = -1: skipped
            }
        }
#pragma unroll
        for (int j = 0; j < MERGE_CPT; ++j) {
            const uint32_t start = ca[j].x, end = ca[j].y, ts = ca[j].z, te = ca[j].w;
            const int32_t tid = (int32_t)cb[j].z;
            if (tid < 0) continue;
            const uint32_t ilen = end - start;                                   // uint32, :161-162
            if (ilen < prm.min_intron || ilen > prm.max_intron) continue;         // junction_qc
            const uint32_t lr = ((start - ts) >= prm.min_anchor ? 1u : 0u) | ((te - end) >= prm.min_anchor ? 2u : 0u);
            const uint32_t sc = cb[j].w & 0xffu;
            const uint32_t proxy = sc == '+' ? 0u : (sc == '-' ? 1u : 2u);         // :186-193
            const unsigned long long ord = (unsigned long long)cb[j].y << 32 | cb[j].x;
            const unsigned long long nfirst = ~ord;
            const unsigned long long last = proxy == 2u ? ((ord >> 16) << 8 | sc) : 0ull;
            bool done = false;
            if (tid == base_tid && ilen < (1u << 28)) {
                const unsigned long long k = (unsigned long long)start << 30 | (unsigned long long)ilen << 2 | proxy;
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
                K128 key{(unsigned long long)start << 32 | end, ((unsigned long long)(uint32_t)(tid + 1)) << 2 | proxy};
                if (!table_upsert(table, mask, key, 1u, ~ts, te, lr, nfirst, last, counters))
                    spill_entry(spill, spill_cap, counters, key, 1u, ~ts, te, lr, nfirst, last);
            }
        }
        __syncthreads();
        // one global upsert per distinct junction of the tile
        for (uint32_t s = t; s < MERGE_SLOTS; s += MERGE_THREADS) {
            const unsigned long long k = sm.key[s];
            if (k == SKEY_EMPTY) continue;
            const uint32_t start = (uint32_t)(k >> 30), ilen = (uint32_t)(k >> 2) & 0x0fffffffu, proxy = (uint32_t)k & 3u;
            K128 key{(unsigned long long)start << 32 | (uint32_t)(start + ilen),
                     ((unsigned long long)(uint32_t)(base_tid + 1)) << 2 | proxy};
            if (!table_upsert(table, mask, key, sm.count[s], sm.nts[s], sm.te[s], sm.lr[s], sm.nfirst[s], sm.last[s], counters))
                spill_entry(spill, spill_cap, counters, key, sm.count[s], sm.nts[s], sm.te[s], sm.lr[s], sm.nfirst[s], sm.last[s]);
        }
        __syncthreads();
    }
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

void launch_junction_merge(const Cand* cands, const uint32_t* d_n_cand, uint32_t n_cand_bound, const ScanParams& p,
                           Slot* table, uint32_t table_mask, Slot* spill_slots, uint32_t spill_cap, uint32_t* d_counters,
                           cudaStream_t stream) {
    if (n_cand_bound == 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(junction_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem));
        attr_set = true;
    }
    uint32_t tiles = (n_cand_bound + MERGE_TILE - 1) / MERGE_TILE;
    uint32_t grid = min(tiles, (uint32_t)(2 * num_sms()));
    junction_merge_kernel<<<grid, MERGE_THREADS, sizeof(MergeSmem), stream>>>(
        cands, d_n_cand, n_cand_bound, p, table, table_mask, spill_slots, spill_cap, d_counters);
}

// Re-inserts every occupied slot of `src` (an old table, or the spill list) into `table`.
__global__ void __launch_bounds__(256)
table_rehash_kernel(const Slot* __restrict__ src, uint32_t n_src, Slot* __restrict__ table, uint32_t mask,
                    uint32_t* __restrict__ counters) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_src; i += gridDim.x * blockDim.x) {
        Slot s = src[i];
        if (s.khi == 0ull) continue;
        if (!table_upsert(table, mask, K128{s.klo, s.khi}, s.count, s.nts, s.te, s.lr, s.nfirst, s.last, counters))
            atomicExch(&counters[CTR_CAND_OVERFLOW], 2u);
    }
}

void launch_table_rehash(const Slot* old_table, uint32_t old_slots, Slot* table, uint32_t table_mask,
                         uint32_t* d_counters, cudaStream_t stream) {
    if (old_slots == 0) return;
    uint32_t grid = min((old_slots + 255u) / 256u, (uint32_t)(8 * num_sms()));
    table_rehash_kernel<<<grid, 256, 0, stream>>>(old_table, old_slots, table, table_mask, d_counters);
}

// ------------------------------------------------------------------------------------------------
// finalize: compaction, first-seen ranking, sort
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
table_compact_kernel(const Slot* __restrict__ table, uint32_t n_slots, OutJunction* __restrict__ out, uint32_t out_cap,
                     uint32_t* __restrict__ d_n_out) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t n_round = (n_slots + stride - 1) / stride * stride;     // keep warps converged for the ballot
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
        Slot s;
        bool occ = false;
        if (i < n_slots) {
            const uint4* p = reinterpret_cast<const uint4*>(table + i);
            uint4 k = p[0];
            occ = (k.z | k.w) != 0u;                                     // khi != 0
            if (occ) {
                uint4 v = p[1], w = p[2];
                s.klo = (unsigned long long)k.y << 32 | k.x; s.khi = (unsigned long long)k.w << 32 | k.z;
                s.count = v.x; s.nts = v.y; s.te = v.z; s.lr = v.w;
                s.nfirst = (unsigned long long)w.y << 32 | w.x; s.last = (unsigned long long)w.w << 32 | w.z;
            }
        }
        const uint32_t m = __ballot_sync(0xffffffffu, occ);
        if (m == 0u) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(d_n_out, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (occ) {
            uint32_t o = base + __popc(m & ((1u << lane) - 1u));
            if (o < out_cap) {
                OutJunction j;
                const uint32_t proxy = (uint32_t)s.khi & 3u;
                j.tid = (int32_t)(uint32_t)(s.khi >> 2) - 1;
                j.start = (uint32_t)(s.klo >> 32); j.end = (uint32_t)s.klo;
                j.ts = ~s.nts; j.te = s.te; j.count = s.count; j.name_index = 0;
                j.strand = proxy == 0u ? '+' : (proxy == 1u ? '-' : (uint8_t)(s.last & 0xffu));
                j.left_ok = s.lr & 1u; j.right_ok = (s.lr >> 1) & 1u; j.pad = 0;
                j.first_ord = ~s.nfirst;
                out[o] = j;
            }
        }
    }
}

void launch_table_compact(const Slot* table, uint32_t n_slots, OutJunction* out, uint32_t out_cap, uint32_t* d_n_out,
                          cudaStream_t stream) {
    uint32_t grid = min((n_slots + 255u) / 256u, (uint32_t)(8 * num_sms()));
    table_compact_kernel<<<grid, 256, 0, stream>>>(table, n_slots, out, out_cap, d_n_out);
}

// key builders / gathers for the two-pass LSD sort
__global__ void fin_keys_first(const OutJunction* __restrict__ e, uint32_t n, unsigned long long* __restrict__ k,
                               uint32_t* __restrict__ v) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { k[i] = e[i].first_ord; v[i] = i; }
}
// after sorting by first_ord: v[r] = entry index with rank r  -> name_index = r+1; key = te<<32|name
__global__ void fin_assign_names(OutJunction* __restrict__ e, uint32_t n, const uint32_t* __restrict__ v,
                                 unsigned long long* __restrict__ k2) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        uint32_t i = v[r];
        e[i].name_index = r + 1u;
        k2[r] = (unsigned long long)e[i].te << 32 | (r + 1u);
    }
}
__global__ void fin_keys_major(const OutJunction* __restrict__ e, uint32_t n, const uint32_t* __restrict__ v,
                               const uint32_t* __restrict__ contig_rank, uint32_t n_contigs,
                               unsigned long long* __restrict__ k3) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) {
        const OutJunction& j = e[v[r]];
        uint32_t cr = (uint32_t)j.tid < n_contigs ? contig_rank[j.tid] : 0x40000000u + (uint32_t)j.tid;
        k3[r] = (unsigned long long)cr << 32 | j.ts;
    }
}
__global__ void fin_gather(const OutJunction* __restrict__ e, uint32_t n, const uint32_t* __restrict__ v,
                           OutJunction* __restrict__ out) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r < n) out[r] = e[v[r]];
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

size_t finalize_sort_workspace_bytes(uint32_t n) {
    size_t cub_bytes = 0;
    cub::DoubleBuffer<unsigned long long> dk(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, dk, dv, (int)n);
    return align_up(cub_bytes, 256) + 2 * align_up((size_t)n * 8, 256) + 2 * align_up((size_t)n * 4, 256) + 256;
}

// entries[0..n) -> scratch[0..n) sorted by (contig_rank, ts, te, name_index); names ranked by first_ord.
void launch_finalize_sort(OutJunction* entries, OutJunction* scratch, uint32_t n, const uint32_t* contig_rank,
                          uint32_t n_contigs, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    if (n == 0) return;
    unsigned char* w = static_cast<unsigned char*>(workspace);
    size_t kb = align_up((size_t)n * 8, 256), vb = align_up((size_t)n * 4, 256);
    unsigned long long* k0 = reinterpret_cast<unsigned long long*>(w); w += kb;
    unsigned long long* k1 = reinterpret_cast<unsigned long long*>(w); w += kb;
    uint32_t* v0 = reinterpret_cast<uint32_t*>(w); w += vb;
    uint32_t* v1 = reinterpret_cast<uint32_t*>(w); w += vb;
    size_t cub_bytes = workspace_bytes - (size_t)(w - static_cast<unsigned char*>(workspace));
    const uint32_t g = (n + 255u) / 256u;

    cub::DoubleBuffer<unsigned long long> dk(k0, k1);
    cub::DoubleBuffer<uint32_t> dv(v0, v1);
    // 1. rank by first_ord
    fin_keys_first<<<g, 256, 0, stream>>>(entries, n, dk.Current(), dv.Current());
    cub::DeviceRadixSort::SortPairs(w, cub_bytes, dk, dv, (int)n, 0, 64, stream);
    // 2. minor key (thick_end, name)
    fin_assign_names<<<g, 256, 0, stream>>>(entries, n, dv.Current(), dk.Alternate());
    dk.selector ^= 1;   // keys now live in the former alternate buffer, values stay current
    cub::DeviceRadixSort::SortPairs(w, cub_bytes, dk, dv, (int)n, 0, 64, stream);
    // 3. major key (contig string rank, thick_start); radix sort is stable
    fin_keys_major<<<g, 256, 0, stream>>>(entries, n, dv.Current(), contig_rank, n_contigs, dk.Alternate());
    dk.selector ^= 1;
    cub::DeviceRadixSort::SortPairs(w, cub_bytes, dk, dv, (int)n, 0, 64, stream);
    fin_gather<<<g, 256, 0, stream>>>(entries, n, dv.Current(), scratch);
}

}  // namespace rtjx
