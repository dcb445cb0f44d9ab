// regtools_b200/csrc/inflate.cu — BGZF/DEFLATE decompression on the device.
//
// Replaces, for whole-file runs, the per-block zlib inflate the reference performs on one host
// thread (/root/reference/src/utils/htslib/bgzf.c:292-316 inflate_block, :421-546 read_block).
// BGZF blocks are independent raw-DEFLATE streams of at most 64 KiB output, so the unit of
// parallelism is the block: one warp per block.  Lane 0 decodes Huffman symbols (the inherently
// serial part) 32 at a time into a shared-memory queue using 10-bit/9-bit lookup tables built per
// DEFLATE block in shared memory; all 32 lanes then place the literals (one coalesced byte store
// per lane) and perform the LZ77 copies cooperatively.  CRC32 is not checked (the reference does
// not check it either).  Any malformed stream sets the block's status to non-zero; the host then
// falls back to its own inflate for the whole run.
#include "jx_device.cuh"
#include <cstdlib>

namespace rtjx {

constexpr int INF_WARPS   = 8;                   // warps (= BGZF blocks in flight) per CTA
constexpr int LIT_BITS    = 10;
constexpr int DIST_BITS   = 9;

struct InflateWarpSmem {
    uint16_t lit_fast[1 << LIT_BITS];            // sym << 4 | len  (0 = needs the slow path)
    uint16_t dist_fast[1 << DIST_BITS];
    uint16_t lit_count[16], dist_count[16];      // canonical decode (slow path): codes per length
    uint16_t lit_sym[288], dist_sym[32];         // symbols sorted by (length, symbol)
    uint8_t  lens[320];                          // code lengths of the current block (lit + dist)
    uint32_t q[32];                              // decoded symbols: len << 16 | dist  or  0x80000000 | byte... see below
};

__constant__ uint16_t c_len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
__constant__ uint8_t c_len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint16_t c_dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
__constant__ uint8_t c_clen_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// LSB-first bit reader over global memory, 32-bit aligned refills (lane 0 only).
struct BitReader {
    const uint32_t* p;      // next aligned word
    uint64_t bb;            // bit buffer
    int nb;                 // valid bits
    __device__ __forceinline__ void init(const uint8_t* src) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        p = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        const int mis = (int)(a & 3);
        bb = (uint64_t)__ldg(p++) >> (8 * mis);
        nb = 32 - 8 * mis;
    }
    __device__ __forceinline__ void need(int n) {         // n <= 32
        if (nb < n) { bb |= (uint64_t)__ldg(p++) << nb; nb += 32; }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { bb >>= n; nb -= n; }
    __device__ __forceinline__ uint32_t get(int n) { need(n); uint32_t v = peek(n); drop(n); return v; }
    // byte position (relative to `base`) of the next unread byte after discarding partial-byte bits
    __device__ __forceinline__ const uint8_t* byte_ptr() const {
        return reinterpret_cast<const uint8_t*>(p) - (nb >> 3);
    }
};

// Word-window bit reader for the symbol loop of the warp-per-block kernel (lane 0): the stream position is
// (w0 | w1 | w2 = three consecutive aligned words, `off` = bit offset into w0).  A 32-bit window at any bit
// offset is ONE funnel shift; consuming bits is one add; a new word is needed once per 32 consumed bits and is
// requested one word ahead so that its latency overlaps the decoding of the current word.
struct WinReader {
    const uint32_t* p;      // address of the word after w2
    uint32_t w0, w1, w2;
    uint32_t off;           // 0..31 (+ consumed, normalised by advance())
    __device__ __forceinline__ void init(const uint8_t* src) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        p = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        off = (uint32_t)(a & 3) * 8;
        w0 = __ldg(p++); w1 = __ldg(p++); w2 = __ldg(p++);
    }
    __device__ __forceinline__ uint32_t window() const { return __funnelshift_r(w0, w1, off); }   // 32 valid bits
    __device__ __forceinline__ void consume(uint32_t n) {                                            // n <= 32
        off += n;
        if (off >= 32u) { off -= 32u; w0 = w1; w1 = w2; w2 = __ldg(p++); }
    }
    __device__ __forceinline__ uint32_t get(uint32_t n) { const uint32_t v = window() & ((1u << n) - 1u); consume(n); return v; }   // n <= 31
    // from / to the generic reader (block headers keep using BitReader)
    __device__ __forceinline__ void from(const BitReader& br) {
        // BitReader state: next unread bit is (nb) bits before the end of the words fetched so far
        const uint8_t* byte = reinterpret_cast<const uint8_t*>(br.p) - ((br.nb + 7) >> 3);
        const uint32_t bit_in_byte = (uint32_t)((8 - (br.nb & 7)) & 7);
        init(byte);
        off += bit_in_byte;
        if (off >= 32u) { off -= 32u; w0 = w1; w1 = w2; w2 = __ldg(p++); }
    }
    __device__ __forceinline__ void to(BitReader& br) const {
        const uint8_t* w0_addr = reinterpret_cast<const uint8_t*>(p - 3);
        br.init(w0_addr + (off >> 3));
        br.drop((int)(off & 7u));
    }
};

// Canonical decode of a code LONGER than the fast table's `fast_bits`: the fast table has already established that no
// code of up to fast_bits bits matches, so the bit-by-bit recurrence starts at length fast_bits + 1 with the state it
// would have reached there — code = the first fast_bits stream bits MSB-first (one brev), first/index = the table's
// constants `pre` (slow_decode_pre).  On worst-case-entropy BAMs (random SEQ/QUAL) ~12 % of the symbols take this path;
// restarting from length 1 made it 61 % of the kernel's instructions (profiles/r1_bgzf_inflate_stall_lines.txt).
struct SlowPre { int first, index; };
__device__ __forceinline__ SlowPre slow_decode_pre(const uint16_t* count, int fast_bits) {
    SlowPre p{0, 0};
    for (int len = 1; len <= fast_bits; ++len) { const int c = count[len]; p.index += c; p.first += c; p.first <<= 1; }
    return p;
}
__device__ __forceinline__ int slow_decode_w(WinReader& wr, const uint16_t* count, const uint16_t* sym, int fast_bits, SlowPre pre) {
    uint32_t win = wr.window();
    int code = (int)(__brev(win) >> (32 - fast_bits)) << 1, first = pre.first, index = pre.index;
    win >>= fast_bits;
    for (int len = fast_bits + 1; len <= 15; ++len) {
        code |= (int)(win & 1u); win >>= 1;
        const int c = count[len];
        if (code - c < first) { wr.consume((uint32_t)len); return sym[index + (code - first)]; }
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

// canonical-code slow decode (one bit at a time), used for codes longer than the fast table
__device__ __forceinline__ int slow_decode(BitReader& br, const uint16_t* count, const uint16_t* sym) {
    int code = 0, first = 0, index = 0;
    for (int len = 1; len <= 15; ++len) {
        code |= (int)br.get(1);
        const int c = count[len];
        if (code - c < first) return sym[index + (code - first)];
        index += c; first += c; first <<= 1; code <<= 1;
    }
    return -1;
}

// Builds count[]/sym[] (canonical order) and the fast table for `n` symbols with lengths `len[]`.
// Executed by the whole warp; returns false (all lanes) on an over-subscribed code.
__device__ bool build_tables(const uint8_t* len, int n, uint16_t* count, uint16_t* sym, uint16_t* fast, int fast_bits,
                             uint32_t lane, uint32_t L, uint32_t gmask) {
    for (int i = lane; i < 16; i += L) count[i] = 0;
    for (int i = lane; i < (1 << fast_bits); i += L) fast[i] = 0;
    __syncwarp(gmask);
    if (lane == 0) {
        for (int i = 0; i < n; ++i) count[len[i]]++;
    }
    __syncwarp(gmask);
    // offsets and over-subscription check (every lane computes the same small loop)
    uint16_t offs[16];
    int left = 1;
    bool ok = true;
    offs[1] = 0;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1; left -= count[l];
        if (left < 0) ok = false;
        if (l < 15) offs[l + 1] = offs[l] + count[l];
    }
    if (!ok) return false;
    // first canonical code of each length
    uint32_t next_code[16];
    {
        uint32_t code = 0;
        next_code[0] = 0;
        for (int l = 1; l <= 15; ++l) { code = (code + (l > 1 ? count[l - 1] : 0)) << 1; next_code[l] = code; }
    }
    if (lane == 0) {
        // symbol table in canonical order + fast-table fill (serial over symbols: <= 288)
        uint16_t o[16];
        for (int l = 0; l < 16; ++l) o[l] = offs[l < 1 ? 1 : l];
        for (int s = 0; s < n; ++s) {
            const int l = len[s];
            if (!l) continue;
            sym[o[l]++] = (uint16_t)s;
            const uint32_t code = next_code[l]++;
            if (l <= fast_bits) {
                const uint32_t rev = __brev(code) >> (32 - l);                 // LSB-first bit order
                const uint16_t e = (uint16_t)(s << 4 | l);
                for (uint32_t k = rev; k < (1u << fast_bits); k += 1u << l) fast[k] = e;
            }
        }
    }
    __syncwarp(gmask);
    return true;
}

struct BgzfBlock {   // same layout as BgzfBlockDesc (jx_device.cuh)
    uint32_t in_off;     // offset of the raw deflate payload inside the compressed buffer
    uint32_t in_len;
    uint32_t out_off;    // offset inside the inflated buffer
    uint32_t out_len;    // ISIZE
};

// L = lanes per decoder (32, 16 or 8).  With L < 32 a warp decodes 32/L BGZF blocks at once: the leaders of the
// lane groups run the (serial) symbol loop in lockstep, so one issued instruction advances 32/L streams, and
// each group of L lanes places its own round of L symbols.  Every warp-level primitive below is restricted to
// the group's lane mask.
template <int L>
__global__ void __launch_bounds__(INF_WARPS * 32)
bgzf_inflate_kernel(const uint8_t* __restrict__ comp, const BgzfBlock* __restrict__ blocks, uint32_t n_blocks,
                    uint8_t* __restrict__ out, uint32_t* __restrict__ status) {
    constexpr int DEC_PER_CTA = INF_WARPS * 32 / L;
    extern __shared__ __align__(16) unsigned char inf_smem_raw[];
    InflateWarpSmem* smem = reinterpret_cast<InflateWarpSmem*>(inf_smem_raw);
    const uint32_t lane = threadIdx.x % L, wib = threadIdx.x / L;               // lane within the group, group within the CTA
    const uint32_t gshift = (threadIdx.x & 31u) / L * L;                          // first warp lane of the group
    const uint32_t gmask = (L == 32 ? 0xffffffffu : ((1u << (L & 31)) - 1u) << gshift);
    const uint32_t b = blockIdx.x * DEC_PER_CTA + wib;
    if (b >= n_blocks) return;
    InflateWarpSmem& sm = smem[wib];
    const BgzfBlock blk = blocks[b];
    uint8_t* dst = out + blk.out_off;
    const uint32_t cap = blk.out_len;
    uint32_t opos = 0;                // bytes produced so far (uniform across the warp)
    uint32_t err = 0;

    BitReader br;                     // only lane 0's copy is meaningful
    if (lane == 0) br.init(comp + blk.in_off);

    for (bool last = false; !last && !err;) {
        // ---- block header (lane 0), broadcast
        uint32_t hdr = 0;
        if (lane == 0) hdr = br.get(3);
        hdr = __shfl_sync(gmask, hdr, 0, L);
        last = hdr & 1u;
        const uint32_t btype = hdr >> 1;
        if (btype == 0) {
            // ---- stored block: LEN, NLEN, raw bytes
            uint32_t len = 0; const uint8_t* src = nullptr;
            if (lane == 0) {
                br.drop(br.nb & 7);                                   // to the next byte boundary
                const uint32_t l = br.get(16), nl = br.get(16);
                if ((l ^ nl) != 0xffffu) len = 0xffffffffu;
                else { len = l; src = br.byte_ptr(); }
            }
            len = __shfl_sync(gmask, len, 0, L);
            if (len == 0xffffffffu || opos + len > cap) { err = 2; break; }
            const unsigned long long s64 = __shfl_sync(gmask, (unsigned long long)reinterpret_cast<uintptr_t>(src), 0, L);
            src = reinterpret_cast<const uint8_t*>((uintptr_t)s64);
            for (uint32_t i = lane; i < len; i += L) dst[opos + i] = __ldg(src + i);
            opos += len;
            if (lane == 0) br.init(src + len);
            continue;
        }
        if (btype == 3) { err = 3; break; }
        // ---- code lengths
        int hlit = 288, hdist = 30;
        if (btype == 1) {
            for (int i = lane; i < 288; i += L) sm.lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
            for (int i = lane; i < 30; i += L) sm.lens[288 + i] = 5;
            __syncwarp(gmask);
        } else {
            uint32_t e2 = 0;
            if (lane == 0) {
                hlit = (int)br.get(5) + 257; hdist = (int)br.get(5) + 1;
                const int hclen = (int)br.get(4) + 4;
                if (hlit > 286 || hdist > 30) e2 = 4;
                uint8_t cl[19];
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (int i = 0; i < hclen; ++i) cl[c_clen_order[i]] = (uint8_t)br.get(3);
                // canonical tables of the code-length code (7-bit max) in registers/local
                uint16_t ccount[8], csym[19];
                for (int i = 0; i < 8; ++i) ccount[i] = 0;
                for (int i = 0; i < 19; ++i) ccount[cl[i]]++;
                uint16_t coffs[8]; coffs[1] = 0;
                for (int l = 1; l < 7; ++l) coffs[l + 1] = coffs[l] + ccount[l];
                for (int i = 0; i < 19; ++i) if (cl[i]) csym[coffs[cl[i]]++] = (uint16_t)i;
                ccount[0] = 0;
                int idx = 0;
                while (idx < hlit + hdist && !e2) {
                    // decode one code-length symbol
                    int code = 0, first = 0, index = 0, symv = -1;
                    for (int l = 1; l <= 7; ++l) {
                        code |= (int)br.get(1);
                        const int c = ccount[l];
                        if (code - c < first) { symv = csym[index + (code - first)]; break; }
                        index += c; first += c; first <<= 1; code <<= 1;
                    }
                    if (symv < 0) { e2 = 5; break; }
                    if (symv < 16) sm.lens[idx++] = (uint8_t)symv;
                    else {
                        int rep, val = 0;
                        if (symv == 16) { if (idx == 0) { e2 = 6; break; } val = sm.lens[idx - 1]; rep = 3 + (int)br.get(2); }
                        else if (symv == 17) rep = 3 + (int)br.get(3);
                        else rep = 11 + (int)br.get(7);
                        if (idx + rep > hlit + hdist) { e2 = 7; break; }
                        while (rep--) sm.lens[idx++] = (uint8_t)val;
                    }
                }
                if (!e2 && sm.lens[256] == 0) e2 = 8;                 // no end-of-block code
                // move the distance lengths to a fixed place (288..)
                if (!e2) {
                    uint8_t tmp[30];
                    for (int i = 0; i < hdist; ++i) tmp[i] = sm.lens[hlit + i];
                    for (int i = hlit; i < 288; ++i) sm.lens[i] = 0;
                    for (int i = 0; i < hdist; ++i) sm.lens[288 + i] = tmp[i];
                    for (int i = hdist; i < 30; ++i) sm.lens[288 + i] = 0;
                }
            }
            e2 = __shfl_sync(gmask, e2, 0, L);
            if (e2) { err = e2; break; }
            hlit = 288; hdist = 30;
            __syncwarp(gmask);
        }
        if (!build_tables(sm.lens, hlit, sm.lit_count, sm.lit_sym, sm.lit_fast, LIT_BITS, lane, L, gmask)) { err = 9; break; }
        // an incomplete distance code with a single symbol is legal; over-subscription is not
        if (!build_tables(sm.lens + 288, hdist, sm.dist_count, sm.dist_sym, sm.dist_fast, DIST_BITS, lane, L, gmask)) { err = 10; break; }

        // ---- symbols, 32 per round
        WinReader wr;
        SlowPre lit_pre{0, 0}, dist_pre{0, 0};
        if (lane == 0) { wr.from(br); lit_pre = slow_decode_pre(sm.lit_count, LIT_BITS); dist_pre = slow_decode_pre(sm.dist_count, DIST_BITS); }
        for (bool eob = false; !eob && !err;) {
            uint32_t n = 0, flag = 0;                 // flag: 1 = end of block seen, 2+ = error
            if (lane == 0) {
                while (n < (uint32_t)L) {
                    uint32_t win = wr.window();
                    int sym;
                    uint32_t e = sm.lit_fast[win & ((1u << LIT_BITS) - 1u)];
                    // Literal fast path: two literals per trip when the next code also sits in the table (the 32-bit window
                    // holds both, 2 x LIT_BITS <= 32), so window / consume / loop overhead is paid once.  e = sym << 4 | len:
                    // 1 <= e <= 0xfff <=> a literal.  Measured 78 -> 74 ms on the 2.1 GB C2 stream.  (Also measured this round,
                    // without effect: 9/8-bit tables for more resident decoders, a shared-memory ring as match source, a
                    // write-combining ring with aligned word stores, deeper bit-stream read-ahead, 8- and 4-lane decoders with
                    // 16 decoders per CTA (126 / 250 ms): the kernel's time follows its warp-instruction count, ~31 per byte.)
                    if (e - 1u < 0xfffu) {
                        const uint32_t l1 = e & 15u;
                        const uint32_t e2 = sm.lit_fast[(win >> l1) & ((1u << LIT_BITS) - 1u)];
                        sm.q[n] = 0x80000000u | (e >> 4);
                        if (e2 - 1u < 0xfffu && n + 1u < (uint32_t)L) {
                            sm.q[n + 1] = 0x80000000u | (e2 >> 4);
                            n += 2; wr.consume(l1 + (e2 & 15u));
                        } else {
                            n += 1; wr.consume(l1);
                        }
                        continue;
                    }
                    if (e) { wr.consume(e & 15u); sym = (int)(e >> 4); }
                    else { sym = slow_decode_w(wr, sm.lit_count, sm.lit_sym, LIT_BITS, lit_pre); if (sym < 0) { flag = 11; break; } }
                    if (sym < 256) { sm.q[n++] = 0x80000000u | (uint32_t)sym; continue; }
                    if (sym == 256) { flag = 1; break; }
                    sym -= 257;
                    if (sym >= 29) { flag = 12; break; }
                    const uint32_t len = c_len_base[sym] + wr.get(c_len_extra[sym]);
                    int ds;
                    win = wr.window();
                    e = sm.dist_fast[win & ((1u << DIST_BITS) - 1u)];
                    if (e) { wr.consume(e & 15u); ds = (int)(e >> 4); }
                    else { ds = slow_decode_w(wr, sm.dist_count, sm.dist_sym, DIST_BITS, dist_pre); if (ds < 0) { flag = 13; break; } }
                    if (ds >= 30) { flag = 14; break; }
                    const uint32_t dist = c_dist_base[ds] + wr.get(c_dist_extra[ds]);
                    sm.q[n++] = len << 16 | dist;       // len <= 258, dist <= 32768
                }
            }
            n = __shfl_sync(gmask, n, 0, L);
            flag = __shfl_sync(gmask, flag, 0, L);
            __syncwarp(gmask);
            if (flag > 1) { err = flag; break; }
            eob = flag == 1;
            // ---- place the round: prefix sum of lengths, literals in parallel, matches in order
            const uint32_t s = lane < n ? sm.q[lane] : 0u;
            const bool is_lit = (s & 0x80000000u) != 0;
            const uint32_t mylen = lane < n ? (is_lit ? 1u : (s >> 16)) : 0u;
            uint32_t x = mylen;
#pragma unroll
            for (int d = 1; d < L; d <<= 1) { uint32_t y = __shfl_up_sync(gmask, x, d, L); if ((int)lane >= d) x += y; }
            const uint32_t total = __shfl_sync(gmask, x, L - 1, L);
            const uint32_t mypos = opos + x - mylen;
            if (opos + total > cap) { err = 15; break; }
            if (lane < n && is_lit) dst[mypos] = (uint8_t)s;
            // A match whose source lies entirely before this round's first byte depends on nothing the round
            // writes: every lane copies its own (short) match at once, one L2 round trip for the whole round.
            // Matches that read bytes produced in this round, and long ones, go through the ordered path below.
            const uint32_t mlen = s >> 16, mdist = s & 0xffffu;
            const bool is_match = lane < n && !is_lit;
            const bool indep = is_match && mlen <= 32u && mypos - opos + (mlen < mdist ? mlen : mdist) <= mdist;
            uint32_t bad = 0;
            if (__ballot_sync(gmask, is_match && mdist > mypos)) bad = 16;
            if (!bad) {
                __syncwarp(gmask);
                if (indep) {
                    const uint8_t* srcp = dst + mypos - mdist;
                    if (mdist >= mlen) { for (uint32_t k = 0; k < mlen; ++k) dst[mypos + k] = __ldcg(srcp + k); }
                    else { for (uint32_t k = 0; k < mlen; ++k) dst[mypos + k] = __ldcg(srcp + (k % mdist)); }
                }
                uint32_t mm = __ballot_sync(gmask, is_match && !indep) >> gshift;
                while (mm) {
                    const int j = __ffs(mm) - 1; mm &= mm - 1;
                    const uint32_t sj = __shfl_sync(gmask, s, j, L);
                    const uint32_t pj = __shfl_sync(gmask, mypos, j, L);
                    const uint32_t len = sj >> 16, dist = sj & 0xffffu;
                    __syncwarp(gmask);                                         // earlier stores of this warp are visible
                    const uint8_t* srcp = dst + pj - dist;
                    if (dist >= len) {
                        for (uint32_t k = lane; k < len; k += L) dst[pj + k] = __ldcg(srcp + k);
                    } else {                                               // overlapping run: periodic with period dist
                        for (uint32_t k = lane; k < len; k += L) dst[pj + k] = __ldcg(srcp + (k % dist));
                    }
                }
            }
            if (bad) { err = bad; break; }
            __syncwarp(gmask);
            opos += total;
        }
        if (lane == 0) wr.to(br);
    }
    if (!err && opos != cap) err = 17;
    if (lane == 0) status[b] = err;
}

// ------------------------------------------------------------------------------------------------
// Lane-per-stream decoder (round 2; the default for launches with enough BGZF blocks)
// ------------------------------------------------------------------------------------------------
// The warp-per-block kernel above issues ~22-31 warp instructions per output byte because its symbol loop runs on one lane.
// Here every LANE decodes its own BGZF block, so one issued instruction advances 32 streams, in two kernels:
//
//  1. bgzf_inflate_lanes_kernel — Huffman decoding.  No lookup tables: a canonical code is decoded by COUNTING, over the 15
//     code lengths, how many left-justified upper bounds `ub[len]` the next 15 stream bits (bit-reversed) reach — thirty
//     compare/add instructions on registers, no branch, no memory access — followed by two small shared-memory reads
//     (adj[len], then the symbol).  A lane therefore needs only 420 bytes of shared memory (symbols in canonical order, one
//     byte each, + per-length adjustments), sixteen warps fit on an SM, and that occupancy — not table size — is what hides the symbol chain's
//     latency.  (First round-2 build: 10-bit lookup tables, 3.2 KB per lane, two warps per SM: 337 ms for the 2.1 GB C2 stream
//     against 74 ms for the kernel above.)
//     One loop, up to two literals and a match (or one DEFLATE block header) per lane per iteration, so lanes reconverge every iteration; zlib
//     cuts DEFLATE blocks after a fixed number of symbols, so the lanes of a warp reach their block headers (the divergent
//     part) in the same iteration.  Literals go straight to their final position (byte stores: combining them into words
//     was measured and cost more issue slots than it saved transactions).
//     Matches are NOT copied here — a copy is a dependent global round trip that would stall all 32 lanes — they are
//     appended to the block's match list (position, length, distance).
//  2. bgzf_match_resolve_kernel — LZ77 copies.  One warp per BGZF block walks the match list 32 matches at a time: matches
//     whose source lies before the round's first byte are copied by their own lanes in parallel, the others (and long ones)
//     cooperatively in order — the placement logic of the warp-per-block kernel, without its decoder.
//
// Error codes are those of the kernel above (+18 input overrun, 19 match list full — not reachable for a block of <= 64 KiB); any non-zero status sends the run to the
// warp-per-block kernel's / host feeder's path.
constexpr int LN_SHARED = 256;                    // per CTA: length / distance base+extra tables (RFC 1951 3.2.5), 64 words
constexpr int LN_STRIDE = 428;                    // bytes of shared memory per lane: 107 words (odd: lanes on distinct banks)
constexpr int LN_LIT_SYM = 0;                     // uint8[288]  low byte of the literal/length symbols in canonical order
constexpr int LN_LIT_AH = 288;                    // uint32[16]  per code length: int16 adj = offset - first_code | uint16 index from which symbols are >= 256
constexpr int LN_DIST_ADJ = 352;                  // int16[16]
constexpr int LN_DIST_SYM = 384;                  // uint8[32]
constexpr int LN_RING = 416;                      // uint32[2]   the bit reader's look-ahead words (cp.async destination)

// LSB-first bit reader of one lane: 64-bit buffer, 32-bit aligned refills.  The next TWO words are always already requested,
// with cp.async into a two-word ring in the lane's shared memory — not into a register: a register that is the destination
// of a load in flight blocks every later instruction of the WARP that touches it, and since the lanes of a warp refill in
// different iterations, a register-staged look-ahead word made each iteration wait for the load the previous one had issued
// (ncu: 1.5 of 6.5 stall cycles per instruction on the long scoreboard).  cp.async.wait_group 1 leaves the newest request in
// flight and waits only for requests at least two refills old.
__device__ __forceinline__ void lane_request(uint32_t slot, const uint32_t* src, bool in_range) {
    if (in_range) {
        asm volatile("{ .reg .u64 g; cvta.to.global.u64 g, %1; cp.async.ca.shared.global [%0], [g], 4; }" :: "r"(slot), "l"(src) : "memory");
    } else {
        asm volatile("st.shared.u32 [%0], %1;" :: "r"(slot), "r"(0u) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
struct LaneBits {
    const uint32_t* p;      // address of the next word to request
    const uint32_t* lim;    // first word that must not be read (end of this block's payload + slack)
    uint64_t bb;
    int nb;
    uint32_t rb, sel;       // shared-space address of the ring; byte offset (0 / 4) of the slot holding the OLDER of the two words
    __device__ __forceinline__ void init(const uint8_t* src, uint32_t len, uint32_t ring) {
        asm volatile("cp.async.wait_all;" ::: "memory");                 // (a re-init after a stored block: nothing of the old stream may still land)
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        p = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
        lim = reinterpret_cast<const uint32_t*>((a + len + 3) & ~(uintptr_t)3) + 2;    // the BGZF trailer (CRC32, ISIZE) follows every payload
        const int mis = (int)(a & 3);
        bb = (uint64_t)__ldg(p++) >> (8 * mis);
        nb = 32 - 8 * mis;
        rb = ring; sel = 0;
        lane_request(rb, p, p < lim); ++p;
        lane_request(rb + 4u, p, p < lim); ++p;
    }
    __device__ __forceinline__ void refill() {                         // afterwards nb >= 33
        if (nb <= 32) {
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            uint32_t w;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(rb + sel) : "memory");
            bb |= (uint64_t)w << nb; nb += 32;
            lane_request(rb + sel, p, p < lim);
            sel ^= 4u;
            ++p;
        }
    }
    __device__ __forceinline__ bool overrun() const { return p > lim + 3; }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)bb & ((1u << n) - 1u); }
    __device__ __forceinline__ void drop(int n) { bb >>= n; nb -= n; }
    __device__ __forceinline__ uint32_t get(int n) { refill(); const uint32_t v = peek(n); drop(n); return v; }   // n <= 16
    // address of the next unread byte after discarding the bits of a partial byte (stored blocks): two words wait in the ring
    __device__ __forceinline__ const uint8_t* byte_ptr() const { return reinterpret_cast<const uint8_t*>(p - 2) - (nb >> 3); }
};

// Canonical code of `n` symbols with lengths len[]: ub[l] (l = 1..15, registers), per-length adjustments and the symbol table
// (shared memory).  Returns false on an over-subscribed code.  ub[l] = (first_code[l] + count[l]) << (15 - l): a 15-bit
// left-justified window v holds a code of length 1 + #{l : v >= ub[l]}; its symbol sits at (v >> (15 - l)) + adj[l] of the
// table, adj[l] = offset[l] - first_code[l].  LIT: symbols are stored as their low byte; within one code length the canonical
// order is ascending, so the symbols >= 256 (end of block, lengths) are the tail of each length's run and one index per
// length (`hi`) tells them apart: ah[l] = uint16(adj[l]) | hi[l] << 16.
template <bool LIT>
__device__ __forceinline__ bool lane_build(const uint8_t* len, int n, uint32_t (&pk)[8], void* adj_out, uint8_t* symtab) {
    uint32_t ub[16];
    uint16_t count[16], offs[16];
    int16_t adj[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { count[i] = 0; adj[i] = 0; }
    for (int i = 0; i < n; ++i) count[len[i]]++;
    int left = 1;
    uint32_t code = 0, off = 0;
    bool ok = true;
#pragma unroll
    for (int l = 1; l <= 15; ++l) {
        const uint32_t c = count[l];
        left <<= 1; left -= (int)c;
        if (left < 0) ok = false;
        ub[l] = (code + c) << (15 - l);
        adj[l] = (int16_t)((int)off - (int)code);
        offs[l] = (uint16_t)off;
        off += c;
        code = (code + c) << 1;
    }
    if (!ok) return false;
    // two bounds per register (lane_code_len): lengths 1..8 in the low halves, 9..15 in the high halves; the 16th slot holds
    // 0x8000, which no 15-bit window reaches
#pragma unroll
    for (int i = 0; i < 8; ++i) pk[i] = ub[i + 1] | (i < 7 ? ub[i + 9] : 0x8000u) << 16;
    const int n_lo = LIT ? (n < 256 ? n : 256) : n;
    for (int s = 0; s < n_lo; ++s) {
        const int l = len[s];
        if (l) symtab[offs[l]++] = (uint8_t)s;
    }
    if (LIT) {
        uint32_t* ah = static_cast<uint32_t*>(adj_out);
        for (int l = 1; l <= 15; ++l) ah[l] = (uint32_t)(uint16_t)adj[l] | (uint32_t)offs[l] << 16;
        for (int s = 256; s < n; ++s) {
            const int l = len[s];
            if (l) symtab[offs[l]++] = (uint8_t)s;
        }
    } else {
        int16_t* a = static_cast<int16_t*>(adj_out);
        for (int l = 1; l <= 15; ++l) a[l] = adj[l];
    }
    return true;
}

// code length of the 15-bit window v (MSB-first): 1 + the number of bounds it reaches; 16 = no code of this table.
// Bounds are <= 0x8000 and v < 0x8000, so (v | 0x8000) - bound stays within its 16-bit half (no borrow between the halves) and
// has bit 15 set exactly when v >= bound: one subtraction compares two lengths, and the fifteen flags are summed by two
// independent accumulators instead of a fifteen-deep chain of predicated increments.
__device__ __forceinline__ uint32_t lane_code_len(uint32_t v, const uint32_t (&pk)[8]) {
    const uint32_t vv = (v | 0x8000u) * 0x10001u;
    uint32_t hi = 0, lo = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t t = vv - pk[i];
        hi += t >> 31;
        lo += t & 0x8000u;
    }
    return 1u + hi + (lo >> 15);
}

// shared-memory loads of the symbol loop by 32-bit shared address: the lane's table base is computed once (a generic pointer
// makes the compiler rebuild the shared window base — S2R + LEA + two IMADs — at every use when registers are tight)
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lds_u8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int32_t lds_s16(uint32_t a) { int32_t v; asm volatile("ld.shared.s16 %0, [%1];" : "=r"(v) : "r"(a)); return v; }

// match records per BGZF block: a match covers >= 3 of the block's <= 65536 bytes, so no valid block has more.  The lists live
// in scratch HBM (175 KB per block, ~1.1 GB per 256 MB group); only the entries a block uses are ever touched.
constexpr uint32_t LN_MATCH_CAP = 21846;

constexpr int LN_THREADS = 64;                    // two warps per CTA: 8 CTAs = 16 warps per SM within the shared-memory budget
__global__ void __launch_bounds__(LN_THREADS, 8)
bgzf_inflate_lanes_kernel(const uint8_t* __restrict__ comp, const BgzfBlock* __restrict__ blocks, uint32_t n_blocks,
                          uint8_t* __restrict__ out, uint32_t* __restrict__ status, uint2* __restrict__ mlist, uint32_t* __restrict__ mcount) {
    extern __shared__ __align__(16) uint8_t lane_smem[];
    const uint32_t b = blockIdx.x * LN_THREADS + threadIdx.x;
    // per-CTA tables: length symbol -> base | extra bits << 16 (29 entries), distance symbol likewise (30 entries, from word 32)
    uint32_t* s_tab = reinterpret_cast<uint32_t*>(lane_smem);
    if (threadIdx.x < 29) { const uint32_t y = threadIdx.x; s_tab[y] = c_len_base[y] | (uint32_t)c_len_extra[y] << 16; }
    if (threadIdx.x < 30) { const uint32_t y = threadIdx.x; s_tab[32 + y] = c_dist_base[y] | (uint32_t)c_dist_extra[y] << 16; }
    __syncthreads();
    uint8_t* T = lane_smem + LN_SHARED + threadIdx.x * LN_STRIDE;
    uint8_t* lit_sym = T + LN_LIT_SYM;
    uint32_t* lit_ah = reinterpret_cast<uint32_t*>(T + LN_LIT_AH);
    int16_t* dist_adj = reinterpret_cast<int16_t*>(T + LN_DIST_ADJ);
    uint8_t* dist_sym = T + LN_DIST_SYM;
    if (b >= n_blocks) return;
    uint32_t sT, sTab;                            // shared-space addresses of the lane's tables and of the CTA's base/extra table
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sT) : "l"(T));
    asm volatile("{ .reg .u64 t; cvta.to.shared.u64 t, %1; cvt.u32.u64 %0, t; }" : "=r"(sTab) : "l"(s_tab));

    const BgzfBlock blk = blocks[b];
    uint8_t* dst = out + blk.out_off;
    const uint32_t cap = blk.out_len;
    uint2* my_matches = mlist + (size_t)b * LN_MATCH_CAP;
    uint32_t opos = 0, err = 0, n_match = 0;
    LaneBits br;
    br.init(comp + blk.in_off, blk.in_len, sT + LN_RING);
    uint32_t ul[8], ud[8];                        // upper bounds per code length of the current DEFLATE block's two codes, two per register
#pragma unroll
    for (int i = 0; i < 8; ++i) { ul[i] = 0; ud[i] = 0; }
    bool in_block = false, last = false;

    for (;;) {
        if (!in_block) {
            // ================= DEFLATE block header (divergent, a few times per BGZF block) =================
            if (last || err) break;
            const uint32_t hdr = br.get(3);
            last = hdr & 1u;
            const uint32_t btype = hdr >> 1;
            if (btype == 0) {                      // stored: LEN, NLEN, raw bytes
                br.drop(br.nb & 7);
                const uint32_t l = br.get(16), nl = br.get(16);
                if ((l ^ nl) != 0xffffu || opos + l > cap) { err = 2; continue; }
                const uint8_t* src = br.byte_ptr();
                if (src + l > reinterpret_cast<const uint8_t*>(br.lim)) { err = 18; continue; }
                for (uint32_t i = 0; i < l; ++i) dst[opos + i] = __ldg(src + i);
                opos += l;
                br.init(src + l, (uint32_t)(reinterpret_cast<const uint8_t*>(br.lim - 2) - (src + l)), sT + LN_RING);
                continue;
            }
            if (btype == 3) { err = 3; continue; }
            uint8_t lens[320];
            if (btype == 1) {
                for (int i = 0; i < 288; ++i) lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                for (int i = 0; i < 30; ++i) lens[288 + i] = 5;
            } else {
                const int hlit = (int)br.get(5) + 257, hdist = (int)br.get(5) + 1, hclen = (int)br.get(4) + 4;
                if (hlit > 286 || hdist > 30) { err = 4; continue; }
                uint8_t cl[19];
#pragma unroll
                for (int i = 0; i < 19; ++i) cl[i] = 0;
                for (int i = 0; i < hclen; ++i) cl[c_clen_order[i]] = (uint8_t)br.get(3);
                uint16_t ccount[8], csym[19], coffs[8];
                for (int i = 0; i < 8; ++i) ccount[i] = 0;
                for (int i = 0; i < 19; ++i) ccount[cl[i]]++;
                coffs[1] = 0;
                for (int l = 1; l < 7; ++l) coffs[l + 1] = coffs[l] + ccount[l];
                for (int i = 0; i < 19; ++i) if (cl[i]) csym[coffs[cl[i]]++] = (uint16_t)i;
                int idx = 0;
                while (idx < hlit + hdist && !err) {
                    br.refill();
                    int code = 0, first = 0, index = 0, symv = -1;
                    for (int l = 1; l <= 7; ++l) {
                        code |= (int)br.peek(1); br.drop(1);
                        const int c = ccount[l];
                        if (code - c < first) { symv = csym[index + (code - first)]; break; }
                        index += c; first += c; first <<= 1; code <<= 1;
                    }
                    if (symv < 0) { err = 5; break; }
                    if (symv < 16) lens[idx++] = (uint8_t)symv;
                    else {
                        int rep, val = 0;
                        if (symv == 16) { if (idx == 0) { err = 6; break; } val = lens[idx - 1]; rep = 3 + (int)br.get(2); }
                        else if (symv == 17) rep = 3 + (int)br.get(3);
                        else rep = 11 + (int)br.get(7);
                        if (idx + rep > hlit + hdist) { err = 7; break; }
                        while (rep--) lens[idx++] = (uint8_t)val;
                    }
                    if (br.overrun()) err = 18;
                }
                if (err) continue;
                if (lens[256] == 0) { err = 8; continue; }
                // distance lengths to their fixed place (288..317); iterate downwards: the ranges may overlap
                for (int i = hdist - 1; i >= 0; --i) lens[288 + i] = lens[hlit + i];
                for (int i = hlit; i < 288; ++i) lens[i] = 0;
                for (int i = hdist; i < 30; ++i) lens[288 + i] = 0;
            }
            if (!lane_build<true>(lens, 288, ul, lit_ah, lit_sym)) { err = 9; continue; }
            if (!lane_build<false>(lens + 288, 30, ud, dist_adj, dist_sym)) { err = 10; continue; }
            in_block = true;
            continue;
        }
        // ================= up to two literal/length symbols, then (if the last one was a length) one match =================
        // Nine tokens in ten of a BAM's DEFLATE streams are literals, and a warp pays for the match path whenever ANY lane has a
        // match: two literal slots per trip through it.  The second code's length is worked out before the first symbol's table
        // lookups have come back (pure register arithmetic on bits already in the buffer: 33 >= 15 + 15 + 3).
        br.refill();
        uint32_t v = __brev((uint32_t)br.bb) >> 17;                    // next 15 bits, first bit most significant
        uint32_t cl = lane_code_len(v, ul);
        if (cl > 15u) { err = 11; in_block = false; continue; }
        uint32_t ah = lds_u32(sT + LN_LIT_AH + 4u * cl);
        br.drop((int)cl);
        const uint32_t v2 = __brev((uint32_t)br.bb) >> 17;
        const uint32_t cl2 = lane_code_len(v2, ul);                    // (meaningless, and unused, if the first symbol is not a literal)
        uint32_t ix = (v >> (15u - cl)) + (uint32_t)(int)(int16_t)(ah & 0xffffu);
        uint32_t sym = lds_u8(sT + LN_LIT_SYM + ix) + (ix >= (ah >> 16) ? 256u : 0u);
        if (sym < 256u) {
            if (opos >= cap) { err = 15; in_block = false; continue; }
            dst[opos++] = (uint8_t)sym;
            if (cl2 > 15u) { err = 11; in_block = false; continue; }
            ah = lds_u32(sT + LN_LIT_AH + 4u * cl2);
            br.drop((int)cl2);
            ix = (v2 >> (15u - cl2)) + (uint32_t)(int)(int16_t)(ah & 0xffffu);
            sym = lds_u8(sT + LN_LIT_SYM + ix) + (ix >= (ah >> 16) ? 256u : 0u);
            if (sym < 256u) {
                if (opos >= cap) { err = 15; in_block = false; continue; }
                dst[opos++] = (uint8_t)sym;
                continue;
            }
        }
        if (sym == 256u) { in_block = false; if (br.overrun()) err = 18; continue; }
        sym -= 257u;
        if (sym >= 29u) { err = 12; in_block = false; continue; }
        // length: base and extra bits from the symbol (RFC 1951 3.2.5)
        br.refill();                                                   // 33 bits: length extra (<= 5) + distance code (<= 15) + distance extra (<= 13)
        const uint32_t lt = lds_u32(sTab + 4u * sym), lx = lt >> 16;
        const uint32_t len = (lt & 0xffffu) + br.peek((int)lx);
        br.drop((int)lx);
        v = __brev((uint32_t)br.bb) >> 17;
        cl = lane_code_len(v, ud);
        if (cl > 15u) { err = 13; in_block = false; continue; }
        const uint32_t ds = lds_u8(sT + LN_DIST_SYM + (uint32_t)((int)(v >> (15u - cl)) + lds_s16(sT + LN_DIST_ADJ + 2u * cl)));
        br.drop((int)cl);
        if (ds >= 30u) { err = 14; in_block = false; continue; }
        const uint32_t dt = lds_u32(sTab + 128u + 4u * ds), dx = dt >> 16;
        const uint32_t dist = (dt & 0xffffu) + br.peek((int)dx);
        br.drop((int)dx);
        if (dist > opos || opos + len > cap || n_match >= LN_MATCH_CAP) {
            err = dist > opos ? 16u : (opos + len > cap ? 15u : 19u); in_block = false; continue;
        }
        my_matches[n_match++] = make_uint2(opos | len << 16, dist);     // the copy itself happens in bgzf_match_resolve_kernel
        opos += len;
    }
    if (!err && opos != cap) err = 17;
    status[b] = err;
    mcount[b] = err ? 0u : n_match;
}

// LZ77 copies of the blocks decoded by bgzf_inflate_lanes_kernel: one warp per BGZF block, 32 matches per round.
constexpr int MR_WARPS = 8;
__global__ void __launch_bounds__(MR_WARPS * 32)
bgzf_match_resolve_kernel(const BgzfBlock* __restrict__ blocks, uint32_t n_blocks, uint8_t* __restrict__ out,
                          const uint2* __restrict__ mlist, const uint32_t* __restrict__ mcount) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t b = blockIdx.x * MR_WARPS + (threadIdx.x >> 5);
    if (b >= n_blocks) return;
    const uint32_t nm = mcount[b];
    if (nm == 0) return;
    uint8_t* dst = out + blocks[b].out_off;
    const uint2* ml = mlist + (size_t)b * LN_MATCH_CAP;
    for (uint32_t r0 = 0; r0 < nm; r0 += 32) {
        const bool have = r0 + lane < nm;
        const uint2 m = have ? ml[r0 + lane] : make_uint2(0u, 1u);
        const uint32_t mypos = m.x & 0xffffu, mlen = m.x >> 16, mdist = m.y;
        const uint32_t round_first = __shfl_sync(0xffffffffu, mypos, 0);
        // A match whose source lies entirely before this round's first byte depends on nothing the round writes
        const bool indep = have && mlen <= 32u && mypos + (mlen < mdist ? mlen : mdist) <= round_first + mdist;
        // (plain loads: every byte of this block is written by this warp — its L1 sees its own stores — or by the kernel before)
        if (indep) {
            // the source bytes all lie before the round's first byte: eight loads in flight, then eight stores
            const uint8_t* srcp = dst + mypos - mdist;
            for (uint32_t k0 = 0; k0 < mlen; k0 += 8) {
                uint8_t t[8];
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j) if (k0 + j < mlen) t[j] = srcp[mdist >= mlen ? k0 + j : (k0 + j) % mdist];
#pragma unroll
                for (uint32_t j = 0; j < 8; ++j) if (k0 + j < mlen) dst[mypos + k0 + j] = t[j];
            }
        }
        uint32_t mm = __ballot_sync(0xffffffffu, have && !indep);
        while (mm) {
            const int j = __ffs(mm) - 1; mm &= mm - 1;
            const uint32_t pj = __shfl_sync(0xffffffffu, mypos, j), len = __shfl_sync(0xffffffffu, mlen, j), dist = __shfl_sync(0xffffffffu, mdist, j);
            __syncwarp();                                          // earlier stores of this warp are visible
            const uint8_t* srcp = dst + pj - dist;
            if (dist >= len) {
                for (uint32_t k = lane; k < len; k += 32) dst[pj + k] = srcp[k];
            } else {                                               // overlapping run: periodic with period dist
                for (uint32_t k = lane; k < len; k += 32) dst[pj + k] = srcp[k % dist];
            }
        }
        __syncwarp();
    }
}

size_t bgzf_inflate_scratch_bytes(uint32_t n_blocks) { return (size_t)n_blocks * LN_MATCH_CAP * sizeof(uint2) + (size_t)n_blocks * 4 + 256; }

static void launch_lanes(const uint8_t* comp, const BgzfBlock* bl, uint32_t n_blocks, uint8_t* out, uint32_t* status, void* scratch, cudaStream_t stream) {
    // RTJX_LANES_SMEM_PAD (developer knob): extra dynamic shared memory per CTA, i.e. fewer resident CTAs per SM
    static const int pad = [] { const char* v = getenv("RTJX_LANES_SMEM_PAD"); return v ? atoi(v) : 0; }();
    const int sh = LN_SHARED + LN_THREADS * LN_STRIDE + pad;
    static bool attr = false;
    if (!attr) {
        attr = true;
        cudaFuncSetAttribute(bgzf_inflate_lanes_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (sh > 48 * 1024) cudaFuncSetAttribute(bgzf_inflate_lanes_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sh);
    }
    uint32_t* mcount = static_cast<uint32_t*>(scratch);
    uint2* mlist = reinterpret_cast<uint2*>(static_cast<uint8_t*>(scratch) + (((size_t)n_blocks * 4 + 255) & ~(size_t)255));
    bgzf_inflate_lanes_kernel<<<(n_blocks + LN_THREADS - 1) / LN_THREADS, LN_THREADS, sh, stream>>>(comp, bl, n_blocks, out, status, mlist, mcount);
    bgzf_match_resolve_kernel<<<(n_blocks + MR_WARPS - 1) / MR_WARPS, MR_WARPS * 32, 0, stream>>>(bl, n_blocks, out, mlist, mcount);
}

__global__ void inflate_status_reduce_kernel(const uint32_t* __restrict__ status, uint32_t n, uint32_t* __restrict__ flags) {
    uint32_t bad = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) bad |= status[i];
    if (__any_sync(0xffffffffu, bad != 0) && (threadIdx.x & 31u) == 0) atomicOr(flags, 8u);   // FEED_FLAG_INFLATE
}
void launch_inflate_status_reduce(const uint32_t* status, uint32_t n_blocks, uint32_t* flags, cudaStream_t stream) {
    if (!n_blocks) return;
    inflate_status_reduce_kernel<<<(n_blocks + 1023) / 1024 > 64 ? 64 : (n_blocks + 1023) / 1024, 256, 0, stream>>>(status, n_blocks, flags);
}

void launch_bgzf_inflate(const uint8_t* comp, const void* blocks, uint32_t n_blocks, uint8_t* out, uint32_t* status,
                         void* scratch, cudaStream_t stream) {
    if (n_blocks == 0) return;
    // RTJX_INFLATE_VARIANT: 0 auto, 1 warp per block, 3 lane per stream.  Read at every launch (a getenv, nanoseconds) so that
    // tests can switch kernels inside one process.  The lane-per-stream pair needs `scratch` (bgzf_inflate_scratch_bytes) and
    // pays once a launch has enough blocks to occupy the lanes.
    int mode = 0;
    { const char* v = getenv("RTJX_INFLATE_VARIANT"); mode = v ? atoi(v) : 0; }
    const BgzfBlock* bl = static_cast<const BgzfBlock*>(blocks);
    if (scratch && (mode == 3 || (mode == 0 && n_blocks >= 1024))) {
        launch_lanes(comp, bl, n_blocks, out, status, scratch, stream);
        return;
    }
    // lanes per decoder: 32 = one BGZF block per warp, 16 / 8 = two / four blocks per warp (RTJX_INFLATE_LANES)
    static int lanes = -1;
    if (lanes < 0) { const char* v = getenv("RTJX_INFLATE_LANES"); lanes = v ? atoi(v) : 16; }
    if (lanes == 8) {
        constexpr int D = INF_WARPS * 32 / 8; const size_t sh = D * sizeof(InflateWarpSmem);
        static bool a8 = false; if (!a8) { cudaFuncSetAttribute(bgzf_inflate_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); a8 = true; }
        bgzf_inflate_kernel<8><<<(n_blocks + D - 1) / D, INF_WARPS * 32, sh, stream>>>(comp, bl, n_blocks, out, status);
    } else if (lanes == 16) {
        constexpr int D = INF_WARPS * 32 / 16; const size_t sh = D * sizeof(InflateWarpSmem);
        static bool a16 = false; if (!a16) { cudaFuncSetAttribute(bgzf_inflate_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); a16 = true; }
        bgzf_inflate_kernel<16><<<(n_blocks + D - 1) / D, INF_WARPS * 32, sh, stream>>>(comp, bl, n_blocks, out, status);
    } else {
        constexpr int D = INF_WARPS; const size_t sh = D * sizeof(InflateWarpSmem);
        static bool a32 = false; if (!a32) { cudaFuncSetAttribute(bgzf_inflate_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh); a32 = true; }
        bgzf_inflate_kernel<32><<<(n_blocks + D - 1) / D, INF_WARPS * 32, sh, stream>>>(comp, bl, n_blocks, out, status);
    }
}

}  // namespace rtjx
