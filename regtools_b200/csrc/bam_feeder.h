// regtools_b200/csrc/bam_feeder.h — native host feeder: BGZF inflate + BAM record split + BAI regions.
//
// Replaces the slice of vendored htslib 1.2.1 that `junctions extract` drives
// (/root/reference/src/utils/htslib: sam.c:399-432 bam_read1, bgzf.c:421-577 block read/inflate,
// hts.c:1708-1819,1924-1964 index query / iteration, hts.c:1517-1624 BAI load).  Differences in
// *how*: the file is mmapped, BGZF blocks are inflated by a pool of workers with reusable zlib
// streams (the reference re-inits zlib per block on one thread), and alignments are never
// materialised as bam1_t — the 16 B + 4 B/op the junction path needs are written straight into
// pinned SoA batches.
#pragma once
#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include <functional>
#include <memory>

namespace rtjx {

struct BamHeader {
    std::vector<std::string> names;     // target_name[tid]
    std::vector<int32_t> lengths;
    uint64_t first_record_voffset = 0;  // virtual offset right after the header
};

struct Chunk64 { uint64_t beg, end; };  // virtual offsets

// Parsed .bai or .csi (hts.c:1569-1624).  A CSI is BGZF-compressed, carries min_shift / depth, stores a loff per bin
// and has no linear index; everything else (bins, chunks, pseudo-bin, query) is the same code with those two numbers.
struct BaiIndex {
    struct Bin { uint32_t bin; uint64_t loff; std::vector<Chunk64> chunks; };
    struct Ref { std::vector<Bin> bins; std::vector<uint64_t> ioffset; const Bin* find(uint32_t bin) const; };
    std::vector<Ref> refs;
    uint64_t n_no_coor = 0;
    int min_shift = 14, n_lvls = 5;               // BAI: fixed; CSI: from the file header
    bool is_csi = false;
    static constexpr uint32_t META_BIN = 37450;   // hts.c:1092 for min_shift 14 / 5 levels
    uint32_t meta_bin() const { return ((1u << (3 * n_lvls + 3)) - 1u) / 7u + 1u; }    // hts.c:1092 META_BIN(idx)

    // hts_idx_load order (hts.c:2009-2042): <bam>.csi, <stem>.csi, <bam>.bai, <stem>.bai.  Returns false if absent/bad.
    static bool load_for_bam(const std::string& bam, BaiIndex* out, bool* csi_present);
    bool load(const std::string& path);
    // HTS_IDX_START offset (hts.c:1721-1731); false => iterator would be NULL.
    bool whole_file_start(uint64_t* voff) const;
    // META_BIN chunk 0 of a contig: [first record, end of last record); false if the contig has no reads.
    bool contig_range(int32_t tid, Chunk64* out) const;
    // hts_itr_query (hts.c:1749-1808): sorted, merged chunk list for [beg,end) on tid.
    std::vector<Chunk64> query(int32_t tid, int64_t beg, int64_t end) const;
};

// The SoA batch the feeder fills (host memory owned by the caller, normally pinned).
struct HostBatch {
    int32_t*  tid = nullptr;
    int32_t*  pos = nullptr;
    uint32_t* meta = nullptr;
    uint32_t* cig_off = nullptr;   // cap_reads + 1
    uint32_t* cigar = nullptr;
    uint32_t* bc = nullptr;        // cap_reads; `-b` mode only: barcode dictionary id of every n_cigar > 1 alignment
    uint32_t  cap_reads = 0, cap_ops = 0;
    uint32_t  n_reads = 0, n_ops = 0, n_junction_ops = 0;
    uint64_t  first_ordinal = 0;
};

// What to iterate (sam_itr_querys semantics, sam.c:678-685 / hts.c:1897-1922).
struct IterSpec {
    enum Kind { WholeFile, NoCoor, Region, Contigs } kind = WholeFile;
    int32_t tid = -1; int64_t beg = 0, end = 0;       // Region
    std::vector<int32_t> contigs;                      // Contigs (sharded whole-file run)
};

struct FeederStats {
    uint64_t reads = 0, cigar_ops = 0, bgzf_blocks = 0, compressed_bytes = 0, inflated_bytes = 0;
    double inflate_s = 0, parse_s = 0, wait_s = 0;
};

class BamFile {
public:
    ~BamFile();
    // false: cannot open / not a BGZF BAM ("Unable to open BAM/SAM file.", junctions_extractor.cc:503-506)
    bool open(const std::string& path, std::string* err);
    const BamHeader& header() const { return hdr_; }
    const uint8_t* data() const { return map_; }
    size_t size() const { return size_; }
    int fd() const { return fd_; }
    // bam_name2id (sam.c:262-277): last duplicate wins; -1 if unknown
    int32_t name2id(const std::string& name) const;
private:
    bool read_header(std::string* err);
    const uint8_t* map_ = nullptr; size_t size_ = 0; int fd_ = -1;
    BamHeader hdr_;
};

// hts_parse_reg + name lookup (hts.c:1877-1922).  false => "Unable to iterate to region within BAM."
bool parse_region(const BamFile& bam, const std::string& region, IterSpec* out);

// Callback interface: the feeder asks for an empty batch, fills it, and hands it back.
struct BatchSink {
    virtual ~BatchSink() {}
    virtual HostBatch* acquire() = 0;                 // blocks until a batch buffer is free
    virtual void submit(HostBatch* b) = 0;            // batch is full (or last)
};

// `-b` single-cell mode: the value of the barcode tag (barcode_tag_ = "CB", junctions_extractor.h:181,192,204) of every
// alignment the reference hands to set_junction_barcode (junctions_extractor.cc:362-374,393-395: all alignments with
// n_cigar > 1), dictionary-encoded in first-seen order.  "?" stands for a missing tag (:370).
struct BarcodeDict {
    std::vector<std::string> names;                       // id -> barcode
    uint64_t missing = 0;                                 // alignments without the tag = WARNING lines of the reference (:371)
    uint64_t bad_type = 0;                                // tag present but neither Z nor H: bam_aux2Z returns NULL (sam.c:1309-1315)
                                                          // and the reference dies in std::string(NULL)
    uint32_t intern(const char* s, size_t n);
    void clear();
    BarcodeDict();
    ~BarcodeDict();
    BarcodeDict(const BarcodeDict&) = delete;
    BarcodeDict& operator=(const BarcodeDict&) = delete;
private:
    struct Map;
    Map* map_;
};

struct FeederOptions {
    int n_threads = 0;            // inflate workers (0 = hardware concurrency)
    bool xs_mode = true;          // scan aux for the strand tag (only for n_cigar > 1)
    char tag[2] = {'X', 'S'};
    BarcodeDict* barcodes = nullptr;   // non-NULL: `-b` mode, HostBatch::bc is filled
    char bc_tag[2] = {'C', 'B'};
};

// Streams every alignment selected by `spec`, in the reference's iteration order, into batches.
// Returns false only on setup errors; malformed/truncated data ends the stream silently, as the
// reference's `while (sam_itr_next(...) >= 0)` does (junctions_extractor.cc:525).
bool feed_alignments(const BamFile& bam, const BaiIndex& idx, const IterSpec& spec, const FeederOptions& opt,
                     BatchSink* sink, FeederStats* stats, std::string* err);

// Walks BGZF block headers from compressed offset `coff` (bgzf.c:348-355,525-546): appends one descriptor per
// block until `max_blocks`, `max_comp_bytes`, the end of file, an empty (ISIZE 0) block or a malformed header.
// Returns the compressed offset after the last block taken; *stop is set when the stream ends there.
struct BgzfBlockInfo { uint64_t coff; uint32_t csize, isize; };
uint64_t scan_bgzf_blocks(const BamFile& bam, uint64_t coff, uint64_t end_coff, size_t max_blocks, uint64_t max_comp_bytes,
                          std::vector<BgzfBlockInfo>* out, bool* stop);

// Same walk over a memory copy of file bytes [base_coff, base_coff + n): stops before the first block that is not
// completely inside the buffer (*partial = true) or at an empty / malformed block (*stop = true).
uint64_t scan_bgzf_blocks_mem(const uint8_t* buf, size_t n, uint64_t base_coff, uint64_t end_coff,
                              std::vector<BgzfBlockInfo>* out, bool* stop, bool* partial, bool* untrusted = nullptr);

// Contigs -> shards: runs of consecutive contigs, balanced on compressed bytes (min-max contiguous partition); and the
// byte ranges of such a run, coalesced into one range per run.
std::vector<int32_t> plan_contig_shards(const BamFile& bam, const BaiIndex& idx, int world);
std::vector<Chunk64> coalesced_contig_ranges(const BaiIndex& idx, const std::vector<int32_t>& contigs);

}  // namespace rtjx
