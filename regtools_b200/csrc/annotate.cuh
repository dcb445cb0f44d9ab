// regtools_b200/csrc/annotate.cuh — device-side layout of `junctions annotate`.
//
// Reference: /root/reference/src/junctions/junctions_annotator.cc — per junction: splice-site 2-mers (:94-114), walk of the
// UCSC bins that can hold an overlapping transcript (:367-388), per transcript overlap_ps / overlap_ns (:128-311).  The
// reference does this with std::map / std::set / std::string per junction on one thread; here the annotation is flat
// integer arrays in HBM and one thread annotates one junction.  All of it is integer / byte work bounded by (random) HBM
// access latency; no tensor cores.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rtjx {

// The GTF as the kernel sees it.  Transcripts are numbered in std::string order of their ids (the iteration order of the
// reference's std::map<string, Transcript>, gtf_parser.cc:153-168), so "sorted by index" == the std::set<string> order of
// AnnotatedJunction::transcripts_overlap.
struct AnnGtfView {
    const unsigned long long* bin_key;    // (chrom id << 32 | bin) ascending: the non-empty bins of chrbin_to_transcripts_
    const uint32_t* bin_off;              // n_bins + 1: range of bin_tx
    const uint32_t* bin_tx;               // transcripts of a bin in id order
    uint32_t n_bins;
    const uint32_t* tx_ex_off;            // n_tx + 1: range of ex_start / ex_end
    const uint8_t*  tx_strand;            // exons[0].strand after the sort: 0 '+', 1 '-', 2 anything else (never matches)
    const uint32_t* ex_start;             // exons in the transcript's sorted order (ascending for '+', descending for '-')
    const uint32_t* ex_end;
    uint32_t n_tx;
};

// Junctions after adjust_junction_ends (:66-81): start = BED start + blockSize[0], end = BED end - blockSize[1] + 1.
struct AnnJunctionView {
    const uint32_t* start; const uint32_t* end;
    const uint8_t*  strand;               // 0 '+', 1 '-', 2 anything else
    const int32_t*  chrom;                // index into the chrom tables below
    uint32_t n;
    // per distinct junction chrom
    const int32_t* c_gtf;                 // chrom id in bin_key, -1 = the GTF has no such seqname
    const unsigned long long* c_goff;     // FASTA: start of the sequence in `genome`
    const unsigned long long* c_glen;     // length, ~0ull = the FASTA has no such sequence
    const uint8_t* genome;                // one byte per base, case preserved
};

// One record per junction.
struct alignas(16) AnnOut {
    uint32_t flags;                       // bit0 known_donor, bit1 known_acceptor, bit2 known_junction, bit3 FASTA lacks the contig
    uint32_t n_acceptors, n_exons, n_donors;      // sizes of the three std::sets (unique elements)
    uint32_t n_tx;                        // transcripts_overlap.size()
    uint32_t tx_off_lo, tx_off_hi;        // where the sorted transcript indices start in `items` (64-bit words)
    uint8_t  ss[6];                       // splice_site: left 2-mer, right 2-mer (already reverse-complemented and swapped for '-')
    uint8_t  ss_n[2];                     // their lengths (a clipped fai_fetch returns 0-2 bases)
};
static_assert(sizeof(AnnOut) == 48, "AnnOut must be 48 bytes");
enum { ANN_KNOWN_DONOR = 1, ANN_KNOWN_ACCEPTOR = 2, ANN_KNOWN_JUNCTION = 4, ANN_NO_CONTIG = 8 };

// counters: [0..1] 64-bit cursor of `items` (total words needed), [2] overflow flag, [3] smallest junction index whose contig
// the FASTA lacks (0xffffffff: none)
enum { ANN_CTR_CURSOR = 0, ANN_CTR_OVERFLOW = 2, ANN_CTR_FIRST_MISSING = 3, ANN_CTR_COUNT = 4 };

// One thread per junction: count pass, reservation in `items`, fill pass, in-place sort + unique.  Asynchronous on `stream`.
void launch_annotate(const AnnGtfView& g, const AnnJunctionView& j, int skip_single_exon, unsigned long long* items,
                     unsigned long long items_cap, AnnOut* out, uint32_t* counters, cudaStream_t stream);

}  // namespace rtjx
