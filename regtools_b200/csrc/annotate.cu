// regtools_b200/csrc/annotate.cu — `junctions annotate` on the device, sm_100a.
//
// Reference (one thread, per junction): junctions_annotator.cc:367-388 walks the 7 UCSC bin levels that can hold a
// transcript overlapping the junction, :333-363 filters on strand, :128-213 / :246-311 (overlap_ps / overlap_ns) walk the
// transcript's exons and fill std::sets of skipped exons / donors / acceptors and three cumulative known_* flags; a
// transcript is reported iff the flags are non-zero when ITS walk ends (`return junction.anchor != "N"`) — which depends on
// the transcripts visited before it, so the visit order (level, bin, transcript id) is part of the result.
//
// Here: one thread per junction walks exactly that order twice — a counting pass, one atomic reservation in a global item
// buffer, a filling pass — then sorts its (tiny) lists in place and counts distinct elements.  The exon walk reads
// ex_start / ex_end sequentially; the bin lookup is one binary search per level over the sorted non-empty bins.
#include "annotate.cuh"

namespace rtjx {

#ifndef RTJX_HOST_EMULATION
namespace {
#else
namespace emul {
#endif

__device__ __constant__ uint32_t ANN_BIN_OFFSETS[7] = {32678 + 4096 + 512 + 64 + 8 + 1, 4096 + 512 + 64 + 8 + 1, 512 + 64 + 8 + 1,
                                                       64 + 8 + 1, 8 + 1, 1, 0};      // bedFile.h:59 (sic: 32678)

struct CountSink {
    uint32_t n_tx = 0, n_ex = 0, n_don = 0, n_acc = 0;
    __device__ __forceinline__ void transcript(uint32_t) { ++n_tx; }
    __device__ __forceinline__ void exon(uint32_t, uint32_t) { ++n_ex; }
    __device__ __forceinline__ void donor(uint32_t) { ++n_don; }
    __device__ __forceinline__ void acceptor(uint32_t) { ++n_acc; }
};
struct FillSink {
    unsigned long long* tx; unsigned long long* ex; unsigned long long* don; unsigned long long* acc;
    uint32_t n_tx = 0, n_ex = 0, n_don = 0, n_acc = 0;
    __device__ __forceinline__ void transcript(uint32_t t) { tx[n_tx++] = t; }
    __device__ __forceinline__ void exon(uint32_t s, uint32_t e) { ex[n_ex++] = (unsigned long long)s << 32 | e; }
    __device__ __forceinline__ void donor(uint32_t p) { don[n_don++] = p; }
    __device__ __forceinline__ void acceptor(uint32_t p) { acc[n_acc++] = p; }
};

// overlap_ps (:128-213) / overlap_ns (:246-311) for transcript t; `flags` are the junction's cumulative known_* bits.
template <class Sink>
__device__ __forceinline__ void visit_transcript(const AnnGtfView& g, uint32_t t, uint32_t js, uint32_t je, uint32_t jstrand,
                                                 bool skip_single, uint32_t& flags, Sink& sink) {
    if (g.tx_strand[t] != jstrand) return;                                  // :343-344
    const uint32_t e0 = g.tx_ex_off[t], n = g.tx_ex_off[t + 1] - e0;
    if (n == 0 || (skip_single && n == 1)) return;                          // :131 / :249
    const uint32_t* __restrict__ xs = g.ex_start + e0;
    const uint32_t* __restrict__ xe = g.ex_end + e0;
    bool started = false;
    if (jstrand == 0u) {
        if (xs[0] > je || xe[n - 1] < js) return;                           // :135-137
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t s = xs[i], e = xe[i];
            if (s > je) break;
            if (e == js && i + 1 < n && xs[i + 1] == je) {                  // :144-150 (i + 1 == n: the reference reads past its vector)
                flags |= ANN_KNOWN_DONOR | ANN_KNOWN_ACCEPTOR | ANN_KNOWN_JUNCTION;
            } else {
                if (!started && e >= js) started = true;
                if (started) {
                    if (s > js && e < je && i > 0 && i < n - 1) sink.exon(s, e);
                    if (e > js && e < je && i < n - 1) sink.donor(e);
                    if (s < je && s > js && i > 0) sink.acceptor(s);
                    if (e == js) flags |= ANN_KNOWN_DONOR;
                    if (s == je) flags |= ANN_KNOWN_ACCEPTOR;
                }
            }
        }
    } else {
        if (xe[0] < js || xs[n - 1] > je) return;                           // :253-256
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t s = xs[i], e = xe[i];
            if (e < js) break;
            if (s == je && i + 1 < n && xe[i + 1] == js) {                  // :263-269
                flags |= ANN_KNOWN_DONOR | ANN_KNOWN_ACCEPTOR | ANN_KNOWN_JUNCTION;
            } else {
                if (!started && s <= je) started = true;
                if (started) {
                    if (s > js && e < je && i > 0 && i < n - 1) sink.exon(s, e);
                    if (e > js && e < je && i < n - 1) sink.acceptor(e);
                    if (s < je && s > js) sink.donor(s);
                    if (e == js) flags |= ANN_KNOWN_ACCEPTOR;
                    if (s == je) flags |= ANN_KNOWN_DONOR;
                }
            }
        }
    }
    if (flags & (ANN_KNOWN_DONOR | ANN_KNOWN_ACCEPTOR | ANN_KNOWN_JUNCTION)) sink.transcript(t);   // annotate_anchor != "N" (:211-212)
}

// annotate_junction_with_gtf (:367-388): levels fine -> coarse, bins ascending, transcripts of a bin in id order.
template <class Sink>
__device__ __forceinline__ uint32_t walk_junction(const AnnGtfView& g, int32_t chrom, uint32_t js, uint32_t je, uint32_t jstrand,
                                                  bool skip_single, Sink& sink) {
    uint32_t flags = 0;
    if (chrom < 0 || jstrand > 1u || g.n_bins == 0) return flags;
    uint32_t sb = js >> 14, eb = (je - 1u) >> 14;
    for (int lvl = 0; lvl < 7; ++lvl) {
        const uint32_t lo = sb + ANN_BIN_OFFSETS[lvl], hi = eb + ANN_BIN_OFFSETS[lvl];
        if (lo <= hi) {
            const unsigned long long klo = (unsigned long long)(uint32_t)chrom << 32 | lo, khi = (unsigned long long)(uint32_t)chrom << 32 | hi;
            uint32_t a = 0, b = g.n_bins;                                    // first non-empty bin >= klo
            while (a < b) { const uint32_t m = (a + b) >> 1; if (g.bin_key[m] < klo) a = m + 1; else b = m; }
            for (; a < g.n_bins && g.bin_key[a] <= khi; ++a)
                for (uint32_t k = g.bin_off[a]; k < g.bin_off[a + 1]; ++k)
                    visit_transcript(g, g.bin_tx[k], js, je, jstrand, skip_single, flags, sink);
        }
        sb >>= 3; eb >>= 3;
    }
    return flags;
}

// in-place shell sort of a thread's own list, then the number of distinct values (= std::set::size())
__device__ __forceinline__ uint32_t sort_unique(unsigned long long* v, uint32_t n, bool compact) {
    if (n < 2) return n;
    uint32_t gap = 1;
    while (gap < n / 3) gap = 3 * gap + 1;
    for (; gap > 0; gap /= 3)
        for (uint32_t i = gap; i < n; ++i) {
            const unsigned long long x = v[i];
            uint32_t k = i;
            for (; k >= gap && v[k - gap] > x; k -= gap) v[k] = v[k - gap];
            v[k] = x;
        }
    uint32_t u = 1;
    for (uint32_t i = 1; i < n; ++i)
        if (v[i] != v[i - 1]) { if (compact) v[u] = v[i]; ++u; }
    return u;
}

__device__ __forceinline__ uint8_t comp_base(uint8_t c) {                    // common.h:59-83
    return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
}

__global__ void __launch_bounds__(128)
annotate_kernel(AnnGtfView g, AnnJunctionView jv, int skip_single, unsigned long long* __restrict__ items, unsigned long long items_cap,
                AnnOut* __restrict__ out, uint32_t* __restrict__ counters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= jv.n) return;
    const uint32_t js = jv.start[i], je = jv.end[i], jstrand = jv.strand[i];
    const int32_t c = jv.chrom[i];
    AnnOut o;
    o.flags = 0; o.n_acceptors = o.n_exons = o.n_donors = o.n_tx = 0; o.tx_off_lo = o.tx_off_hi = 0;
    for (int k = 0; k < 6; ++k) o.ss[k] = 0;
    o.ss_n[0] = o.ss_n[1] = 0;

    // ---- get_splice_site (:94-114): fai_fetch of [start+1, start+2] and [end-2, end-1], 1-based inclusive, clipped
    const unsigned long long glen = c >= 0 ? jv.c_glen[c] : ~0ull;
    if (glen == ~0ull) {
        o.flags |= ANN_NO_CONTIG;
        atomicMin(&counters[ANN_CTR_FIRST_MISSING], i);
    } else {
        const uint8_t* gq = jv.genome + jv.c_goff[c];
        uint8_t s1[2] = {0, 0}, s2[2] = {0, 0};
        uint32_t n1 = 0, n2 = 0;
        // fai_fetch (faidx.c:341-415) reads the two numbers of "chrom:a-b" back with atoi into ints (the annotator printed them
        // with uint32 arithmetic, so end - 2 of a junction ending at 1 arrives as -1), decrements a positive beg, clips both to the
        // length and never clamps a negative beg: it then starts that many bytes before the sequence — for -1 the header's
        // newline, skipped as non-graph — and returns the first (end - beg) bases.  (beg < -1 would read header text: as -1.)
        auto fetch = [&](uint32_t a1, uint32_t b1, uint8_t* dst) -> uint32_t {
            long long b = (long long)(int32_t)a1, e = (long long)(int32_t)b1;
            const long long len = (long long)glen;
            if (b > 0) --b;
            if (b >= len) b = len;
            if (e >= len) e = len;
            if (b > e) b = e;
            const long long from = b < 0 ? 0 : b;
            long long n = e - b;
            if (n > len - from) n = len - from;
            if (n > 2) n = 2;
            for (long long k = 0; k < n; ++k) dst[k] = gq[from + k];
            return (uint32_t)(n < 0 ? 0 : n);
        };
        n1 = fetch(js + 1u, js + 2u, s1);
        n2 = fetch(je - 2u, je - 1u, s2);
        if (jstrand == 1u) {                                                   // rev_comp both, print seq2-seq1 (:106-110)
            for (uint32_t k = 0; k < n2; ++k) o.ss[k] = comp_base(s2[n2 - 1 - k]);
            for (uint32_t k = 0; k < n1; ++k) o.ss[3 + k] = comp_base(s1[n1 - 1 - k]);
            o.ss_n[0] = (uint8_t)n2; o.ss_n[1] = (uint8_t)n1;
        } else {
            for (uint32_t k = 0; k < n1; ++k) o.ss[k] = s1[k];
            for (uint32_t k = 0; k < n2; ++k) o.ss[3 + k] = s2[k];
            o.ss_n[0] = (uint8_t)n1; o.ss_n[1] = (uint8_t)n2;
        }
    }

    // ---- pass 1: how many items this junction produces
    const int32_t gc = c >= 0 ? jv.c_gtf[c] : -1;
    CountSink cs;
    const uint32_t flags = walk_junction(g, gc, js, je, jstrand, skip_single != 0, cs);
    o.flags |= flags;
    const unsigned long long total = (unsigned long long)cs.n_tx + cs.n_ex + cs.n_don + cs.n_acc;
    if (total) {
        const unsigned long long base = atomicAdd(reinterpret_cast<unsigned long long*>(counters + ANN_CTR_CURSOR), total);
        if (base + total > items_cap) {
            atomicExch(&counters[ANN_CTR_OVERFLOW], 1u);                       // the host re-runs with the exact size
        } else {
            // ---- pass 2: same walk, items stored; then std::set semantics by sort + unique
            FillSink fs;
            fs.tx = items + base; fs.ex = fs.tx + cs.n_tx; fs.don = fs.ex + cs.n_ex; fs.acc = fs.don + cs.n_don;
            walk_junction(g, gc, js, je, jstrand, skip_single != 0, fs);
            o.n_tx = sort_unique(fs.tx, fs.n_tx, true);
            o.n_exons = sort_unique(fs.ex, fs.n_ex, false);
            o.n_donors = sort_unique(fs.don, fs.n_don, false);
            o.n_acceptors = sort_unique(fs.acc, fs.n_acc, false);
            o.tx_off_lo = (uint32_t)base; o.tx_off_hi = (uint32_t)(base >> 32);
        }
    }
    out[i] = o;
}

}  // namespace
#ifdef RTJX_HOST_EMULATION
using namespace emul;
#endif

#ifndef RTJX_HOST_EMULATION   // tests/emul/annotate_emul.cc runs the kernel body in a host loop to check its logic without a GPU
void launch_annotate(const AnnGtfView& g, const AnnJunctionView& j, int skip_single_exon, unsigned long long* items,
                     unsigned long long items_cap, AnnOut* out, uint32_t* counters, cudaStream_t stream) {
    if (j.n == 0) return;
    annotate_kernel<<<(j.n + 127u) / 128u, 128, 0, stream>>>(g, j, skip_single_exon, items, items_cap, out, counters);
}
#endif

}  // namespace rtjx
