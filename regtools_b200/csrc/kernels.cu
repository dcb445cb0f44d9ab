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
// Two cigar_scan kernels are kept: the block-per-tile kernel (default) and the warp-pipelined persistent kernel (opt-in,
// rtjx_params.scan_variant = 8); the round-1 A/B builds (variants 1, 4, 6, 7 and the probe kernels) are gone — their
// measurements live in profiles/r1_scan_ab.md.
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

static int num_sms();

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
    Smem& sm; Cand* __restrict__ out; uint32_t cap; uint32_t* counters;
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
            if (vr.end[i] > rp) push(a, make_uint4(b.x, b.y, b.z, vr.tag ? strand | (i + 1u) << 8 : strand));
        }
    }
};

// MINB: resident blocks per SM asked of ptxas (caps the registers per thread at 65536 / (S5_THREADS * MINB))
template <int S5_THREADS, int S5_SLAB, int S5_OUT, bool MOTIF = false, bool VREG = false, bool BC = false, int MINB = 1>
__global__ void __launch_bounds__(S5_THREADS, MINB)
cigar_scan_small_kernel(BatchView b, ScanParams prm, Cand* __restrict__ out, uint32_t cap, uint32_t* __restrict__ counters) {
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
    cp_async_wait_all();
    __syncthreads();
    // ---- CIGAR slab of the tile: its address is the one data-dependent address of the path
    const uint32_t lo = sm.off[0], hi = sm.off[n_tile];
    const uint32_t a0 = lo & ~3u;
    // words of the tile's slab staged in shared memory (from a0): all of it unless the tile is denser than S5_SLAB or
    // touches the ragged end of the array; alignments outside the staged window read their ops from global memory
    uint32_t n_st = 0;
    if (hi > lo) {
        const uint32_t end = min(min((hi + 3u) & ~3u, a0 + (uint32_t)S5_SLAB), b.n_ops & ~3u);
        n_st = end > a0 ? end - a0 : 0u;
        const uint32_t nv = n_st >> 2;                         // 16-byte vectors: at most S5_SLAB / 4, i.e. a fixed few per thread
        const uint32_t* src = b.cigar + a0 + 4 * t;
#pragma unroll
        for (int j = 0; j < (S5_SLAB / 4 + S5_THREADS - 1) / S5_THREADS; ++j)
            if (t + j * S5_THREADS < nv) cp_async16(&sm.slab[4 * (t + j * S5_THREADS)], src + 4 * j * S5_THREADS);
    }
    cp_async_commit();
    {   // ---- while the slab is in flight: list the alignments with more than one CIGAR op (junctions_extractor.cc:379);
        // ballots, no scan: the order of the list does not matter
        const uint4 o = *reinterpret_cast<const uint4*>(&sm.off[4 * t]);
        const uint32_t o4 = sm.off[4 * t + 4];
        const uint32_t r0 = 4 * t, lt = (1u << lane) - 1u;
        const bool f0 = r0 + 0 < n_tile && o.y - o.x > 1u, f1 = r0 + 1 < n_tile && o.z - o.y > 1u;
        const bool f2 = r0 + 2 < n_tile && o.w - o.z > 1u, f3 = r0 + 3 < n_tile && o4 - o.w > 1u;
        const uint32_t m0 = __ballot_sync(0xffffffffu, f0), m1 = __ballot_sync(0xffffffffu, f1);
        const uint32_t m2 = __ballot_sync(0xffffffffu, f2), m3 = __ballot_sync(0xffffffffu, f3);
        const uint32_t s1 = __popc(m0), s2 = s1 + __popc(m1), s3 = s2 + __popc(m2), tot = s3 + __popc(m3);
        uint32_t wbase = 0;
        if (lane == 0 && tot) wbase = atomicAdd(&sm.n_work, tot);
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (f0) sm.work[wbase + __popc(m0 & lt)] = (uint16_t)(r0 + 0);
        if (f1) sm.work[wbase + s1 + __popc(m1 & lt)] = (uint16_t)(r0 + 1);
        if (f2) sm.work[wbase + s2 + __popc(m2 & lt)] = (uint16_t)(r0 + 2);
        if (f3) sm.work[wbase + s3 + __popc(m3 & lt)] = (uint16_t)(r0 + 3);
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- walk the listed alignments, one per thread, in rounds of S5_THREADS
    const uint32_t n_work = sm.n_work;
    uint32_t jstrand = 0;                                      // j1.strand == "" before an alignment's first junction
    int32_t rspan[2] = {0, 0};
    const S5Emit<Smem, S5_OUT, MOTIF, VREG, BC> emit{sm, out, cap, counters, &prm, &jstrand, rspan};
    auto flush = [&]() {                                        // warp 0: staged candidates -> HBM, one reservation
        const uint32_t n_out = min(sm.n_out, (uint32_t)S5_OUT);
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
        if (t < 32) flush();
        if (w0 + S5_THREADS < n_work) __syncthreads();
    }
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
            if (vr.end[i] > rp) push(a, make_uint4(b.x, b.y, b.z, vr.tag ? strand | (i + 1u) << 8 : strand));
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

// Element-wise fallback for caller-owned device arrays that are not 16-byte aligned (rtjx_scan_batch, RTJX_LOC_DEVICE): one
// thread per alignment, ops read from global memory, one slot reservation per candidate.  Plain mode only.
struct DirectEmit {
    Cand* __restrict__ out; uint32_t cap; uint32_t* counters;
    __device__ __forceinline__ void operator()(uint32_t start, uint32_t end, uint32_t left, uint32_t right, uint64_t ord, int32_t tid,
                                               uint32_t strand) const {
        store_cand(out, cap, counters, atomicAdd(&counters[CTR_NCAND], 1u), make_uint4(start, end, start - left, end + right),
                   make_uint4((uint32_t)ord, (uint32_t)(ord >> 32), (uint32_t)tid, strand));
    }
};
__global__ void __launch_bounds__(256)
cigar_scan_unaligned_kernel(BatchView b, ScanParams prm, Cand* __restrict__ out, uint32_t cap, uint32_t* __restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b.n_reads) return;
    const uint32_t o0 = b.cig_off[i], n = b.cig_off[i + 1] - o0;
    const int32_t tid = b.tid[i];
    if (n <= 1u || tid < 0) return;                           // junctions_extractor.cc:379
    walk_fast<false>(b.cigar + o0, n, (uint32_t)b.pos[i], tid, read_strand(b.meta[i], prm.strandness), b.first_ordinal + i,
                     DirectEmit{out, cap, counters});
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

void launch_cigar_scan(const BatchView& b, const ScanParams& p, Cand* cands, uint32_t cand_cap, uint32_t* d_counters, cudaStream_t stream) {
    if (b.n_reads == 0) return;
    const uintptr_t align = reinterpret_cast<uintptr_t>(b.tid) | reinterpret_cast<uintptr_t>(b.pos) |
                            reinterpret_cast<uintptr_t>(b.meta) | reinterpret_cast<uintptr_t>(b.cig_off) |
                            reinterpret_cast<uintptr_t>(b.cigar);
    const bool special = p.genome || p.vr.n || b.bc;
    if ((align & 15u) == 0 && p.variant == 8) {               // warp-pipelined kernel (opt-in: rtjx_params.scan_variant = 8)
        if (special) { launch_pipe<384, 3, 8, 2, true>(b, p, cands, cand_cap, d_counters, stream); return; }
        switch (p.cfg) {                   // ring / occupancy configurations measured in profiles/r2_scan_ab_*.json
        case 1: launch_pipe<384, 3, 4, 4>(b, p, cands, cand_cap, d_counters, stream); break;      // 16 warps/SM, 2 tiles in flight each
        case 2: launch_pipe<256, 2, 8, 4>(b, p, cands, cand_cap, d_counters, stream); break;      // 32 warps/SM, 1 in flight
        case 3: launch_pipe<256, 2, 4, 7>(b, p, cands, cand_cap, d_counters, stream); break;      // 28 warps/SM, 1 in flight
        default: launch_pipe<384, 2, 8, 3>(b, p, cands, cand_cap, d_counters, stream); break;     // 24 warps/SM, 1 in flight
        }
        return;
    }
    // block-per-tile kernel (the default): 16-byte aligned arrays; anything else takes the element-wise fallback below
    const uint32_t tiles = (b.n_reads + 511u) / 512u;
    if ((align & 15u) == 0) {
        if (b.bc && p.genome) cigar_scan_small_kernel<128, 1024, 192, true, false, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters);
        else if (b.bc) cigar_scan_small_kernel<128, 1024, 192, false, false, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters);
        else if (p.vr.n && p.genome) cigar_scan_small_kernel<128, 1024, 192, true, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters);
        else if (p.vr.n) cigar_scan_small_kernel<128, 1024, 192, false, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters);
        else if (p.genome) cigar_scan_small_kernel<128, 1024, 192, true><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters);
        // plain mode: occupancy configurations (profiles/r2_scan_ab_block_tiled_*.json; an L2 prefetch of later tiles' columns was
        // measured too and changes nothing: profiles/r2_scan_ab_l2_prefetch_*.json)
        else switch (p.cfg) {
            case 1: cigar_scan_small_kernel<128, 1024, 128, false, false, false, 12><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
            case 2: cigar_scan_small_kernel<128, 1024, 96, false, false, false, 14><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
            case 3: cigar_scan_small_kernel<128, 768, 96, false, false, false, 16><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
            case 4: cigar_scan_small_kernel<128, 1024, 192, false, false, false, 12><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
            case 5: cigar_scan_small_kernel<128, 1024, 192><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
            default: cigar_scan_small_kernel<128, 1024, 128, false, false, false, 12><<<tiles, 128, 0, stream>>>(b, p, cands, cand_cap, d_counters); break;
        }
        return;
    }
    cigar_scan_unaligned_kernel<<<(b.n_reads + 255u) / 256u, 256, 0, stream>>>(b, p, cands, cand_cap, d_counters);
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

struct MergeSmem {
    unsigned long long key[MERGE_SLOTS];
    unsigned long long nfirst[MERGE_SLOTS];
    unsigned long long last[MERGE_SLOTS];
    uint32_t count[MERGE_SLOTS], nts[MERGE_SLOTS], te[MERGE_SLOTS], lr[MERGE_SLOTS];
    unsigned long long pal[MERGE_PAL];           // region << 32 | tid of the (contig, region) pairs the tile's shared-memory table knows
};

__global__ void __launch_bounds__(MERGE_THREADS, 2)
junction_merge_kernel(const Cand* __restrict__ cands, const uint32_t* __restrict__ d_n_cand, uint32_t n_bound,
                      ScanParams prm, TableRef tb, Slot* __restrict__ spill, uint32_t spill_cap,
                      uint32_t* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MergeSmem& sm = *reinterpret_cast<MergeSmem*>(smem_raw);
    const uint32_t t = threadIdx.x;
    const uint32_t n = d_n_cand ? min(*d_n_cand, n_bound) : n_bound;
    const uint32_t n_tiles = (n + MERGE_TILE - 1) / MERGE_TILE;
    uint32_t n_valid = 0;                                       // candidates seen (tid >= 0: not the padding of a scan kernel's chunks)

    for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint32_t tb0 = tile * MERGE_TILE;
        const uint32_t m = min((uint32_t)MERGE_TILE, n - tb0);
        for (uint32_t s = t; s < MERGE_SLOTS; s += MERGE_THREADS) {
            sm.key[s] = SKEY_EMPTY; sm.nfirst[s] = 0ull; sm.last[s] = 0ull;
            sm.count[s] = 0u; sm.nts[s] = 0u; sm.te[s] = 0u; sm.lr[s] = 0u;
        }
        if (t < MERGE_PAL) sm.pal[t] = SKEY_EMPTY;
        // all loads of the tile first (two 128-bit loads per candidate)
        uint4 ca[MERGE_CPT], cb[MERGE_CPT];
#pragma unroll
        for (int j = 0; j < MERGE_CPT; ++j) {
            const uint32_t i = t + j * MERGE_THREADS;
            if (i < m) {
                const uint4* p = reinterpret_cast<const uint4*>(cands + (size_t)tb0 + i);
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
            const uint32_t sc = cb[j].w & 0xffu, vreg = cb[j].w >> 8;           // vreg = variant region + 1 / barcode id + 1 (0 outside those modes)
            const uint32_t proxy = sc == '+' ? 0u : (sc == '-' ? 1u : 2u);         // :186-193
            const unsigned long long ord = (unsigned long long)cb[j].y << 32 | cb[j].x;
            const unsigned long long nfirst = ~ord;
            const unsigned long long last = proxy == 2u ? ((ord >> 16) << 8 | sc) : 0ull;
            bool done = false;
            // (contig, region) -> palette index: the first MERGE_PAL distinct pairs of the tile (a BAM is sorted: nearly always 1-2)
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
        // one global upsert per distinct junction of the tile
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

void launch_junction_merge(const Cand* cands, const uint32_t* d_n_cand, uint32_t n_cand_bound, const ScanParams& p, const TableRef& tb,
                           Slot* spill_slots, uint32_t spill_cap, uint32_t* d_counters, cudaStream_t stream) {
    if (n_cand_bound == 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(junction_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(MergeSmem));
        attr_set = true;
    }
    const uint32_t tiles = (n_cand_bound + MERGE_TILE - 1) / MERGE_TILE;
    const uint32_t grid = max(1u, min(tiles, (uint32_t)(2 * num_sms())));
    junction_merge_kernel<<<grid, MERGE_THREADS, sizeof(MergeSmem), stream>>>(cands, d_n_cand, n_cand_bound, p, tb, spill_slots, spill_cap, d_counters);
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
// merged table of contig shards: ordinals restart on every rank, but a contig lives on one rank and contigs follow the file
struct ByContigFirstOrd {
    __device__ __forceinline__ bool operator()(const OutJunction& a, const OutJunction& b) const {
        return a.tid != b.tid ? a.tid < b.tid : a.first_ord < b.first_ord;
    }
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
    size_t a = 0, b = 0, c = 0;
    cub::DeviceMergeSort::SortKeys(nullptr, c, (OutJunction*)nullptr, (int)n, ByContigFirstOrd());
    cub::DeviceMergeSort::SortKeys(nullptr, a, (OutJunction*)nullptr, (int)n, ByFirstOrd());
    if (c > a) a = c;
    cub::DeviceMergeSort::SortKeys(nullptr, b, (OutJunction*)nullptr, (int)n, ByBedOrder{nullptr, 0});
    return (a > b ? a : b) + 256;
}

// entries[0..n): ranked by first_ord (name_index), then sorted in place by (contig string, ts, te, name).
void launch_finalize_sort(OutJunction* entries, uint32_t n, const uint32_t* contig_rank, uint32_t n_contigs,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream, bool rank_by_contig) {
    if (n == 0) return;
    if (rank_by_contig) cub::DeviceMergeSort::SortKeys(workspace, workspace_bytes, entries, (int)n, ByContigFirstOrd(), stream);
    else cub::DeviceMergeSort::SortKeys(workspace, workspace_bytes, entries, (int)n, ByFirstOrd(), stream);
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
