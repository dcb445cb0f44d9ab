// regtools_b200/csrc/junctions_extractor.h — C++ drop-in shim for the reference class.
//
// Same class name, method names, argument meaning and exception behaviour as
// /root/reference/src/junctions/junctions_extractor.h:39-248 so that the reference's callers
// (src/junctions/junctions_main.cc:45-59, src/cis-splice-effects/cis_splice_effects_identifier.cc:288-290,
// tests/lib/junctions/test_junctions_extractor.cc) compile against it unchanged.  Every method
// forwards to the C ABI (include/rtjx.h); a non-zero status is rethrown as std::runtime_error with
// the reference's message text.  The heavy lifting (BGZF feeder, CUDA kernels) is in
// libregtools_jx.so; this header + junctions_extractor.cc add no compute.
#ifndef RTJX_JUNCTIONS_EXTRACTOR_H_
#define RTJX_JUNCTIONS_EXTRACTOR_H_

#include <stdint.h>
// (the standard headers the reference's header pulls in — its callers lean on them transitively, e.g. setw in
// cis_splice_effects_identifier.cc:251)
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/rtjx.h"

// The reference header opens namespace std for everything that includes it (junctions_extractor.h:36) and its callers rely
// on that (tests/lib/junctions/test_junctions_extractor.cc uses an unqualified `string`): kept, for source compatibility.
using namespace std;

#ifndef BEDFILE_H   // when built inside the reference tree, bedFile.h supplies BED and CHRPOS
typedef uint32_t CHRPOS;                               // src/utils/bedtools/bedFile/bedFile.h:40
struct BED {                                           // the fields of bedFile.h:79-184 the path uses
    std::string chrom;
    CHRPOS start, end;
    std::string name, score, strand;
    std::vector<std::string> fields;
    BED() : start(0), end(0) {}
};
#endif

namespace common {
#ifndef COMMON_H_
// src/utils/common.h:96-101
class cmdline_help_exception : public std::runtime_error {
public:
    explicit cmdline_help_exception(std::string const& msg) : std::runtime_error(msg) {}
};
#endif
}  // namespace common

// junctions_extractor.h:39-112 (the barcodes map stays in the engine: rtjx_write_barcodes prints it)
struct Junction : BED {
    unsigned int read_count;
    CHRPOS thick_start, thick_end;
    bool added;
    bool has_left_min_anchor, has_right_min_anchor;
    std::string color;
    int nblocks;
    // `-b` single cell: barcode -> count (junctions_extractor.h:57-58).  The engine keeps the barcode lists of a run on its side
    // (rtjx_write_barcodes prints them); this member exists for the reference's callers and for junctions built by hand.
    std::unordered_map<std::string, int> barcodes;
    Junction() : read_count(0), thick_start(0), thick_end(0), added(false), has_left_min_anchor(false),
                 has_right_min_anchor(false), color("255,0,0"), nblocks(2) { name = "NA"; }
    Junction(std::string chrom1, CHRPOS start1, CHRPOS end1, CHRPOS thick_start1, CHRPOS thick_end1, std::string strand1)
        : read_count(0), thick_start(thick_start1), thick_end(thick_end1), added(false), has_left_min_anchor(false),
          has_right_min_anchor(false), color("255,0,0"), nblocks(2) {
        chrom = chrom1; start = start1; end = end1; name = "NA"; strand = strand1;
    }
    void print(std::ostream& out) const {              // junctions_extractor.h:90-98
        out << chrom << "\t" << thick_start << "\t" << thick_end << "\t" << name << "\t" << read_count << "\t" << strand
            << "\t" << thick_start << "\t" << thick_end << "\t" << color << "\t" << nblocks
            << "\t" << start - thick_start << "," << thick_end - end << "\t" << "0," << end - thick_start << std::endl;
    }
    void print_barcodes(std::ostream& out) const {     // junctions_extractor.h:99-111
        out << barcodes.size() << "\t";
        for (std::unordered_map<std::string, int>::const_iterator it = barcodes.begin(); it != barcodes.end(); it++) {
            if (it != barcodes.begin()) out << ",";
            out << it->first << ":" << it->second;
        }
        out << std::endl;
    }
};

// junctions_extractor.h:117-146
static inline bool compare_junctions(const Junction& j1, const Junction& j2) {
    if (j1.chrom < j2.chrom) return true;
    if (j1.chrom > j2.chrom) return false;
    if (j1.thick_start < j2.thick_start) return true;
    if (j1.thick_start > j2.thick_start) return false;
    if (j1.thick_end < j2.thick_end) return true;
    if (j1.thick_end > j2.thick_end) return false;
    return j1.name < j2.name;
}

class JunctionsExtractor {
public:
    JunctionsExtractor();                                                   // junctions_extractor.h:184-198
    // 8-arg ctor of cis-splice-effects, including its min_intron := min_anchor initialiser (:199-205)
    JunctionsExtractor(std::string bam1, std::string region1, int strandness1, std::string strand_tag1,
                       uint32_t min_anchor_length1, uint32_t min_intron_length1, uint32_t max_intron_length1,
                       std::string ref1);
    ~JunctionsExtractor();
    JunctionsExtractor(const JunctionsExtractor&) = delete;
    JunctionsExtractor& operator=(const JunctionsExtractor&) = delete;

    std::string get_new_junction_name();                                    // junctions_extractor.cc:152-157
    int parse_options(int argc, char* argv[]);                              // :42-122
    int usage(std::ostream& out = std::cerr);                               // :125-143
    int identify_junctions_from_BAM();                                      // :500-535
    void print_all_junctions(std::ostream& out = std::cout);                // :249-280
    std::vector<Junction> get_all_junctions();                              // :238-246
    std::string get_bam();                                                  // :146-148
    int add_junction(Junction j1);                                          // :174-235

    // Not in the reference: the per-variant loop of cis_splice_effects_identifier.cc:267-311 in one pass over the BAM.
    // Element i is what `JunctionsExtractor(bam, regions[i], ...)` + identify + get_all_junctions() returns; the object
    // must have been constructed with region "." (rtjx_run_regions).
    std::vector<std::vector<Junction> > get_all_junctions_in_regions(const std::vector<std::string>& regions);

    // B200-specific knobs (not in the reference): CUDA device, host inflate threads, contig shard
    void set_device(int device) { device_ = device; }
    void set_threads(int n) { n_threads_ = n; }
    void set_shard(int rank, int world) { shard_rank_ = rank; shard_world_ = world; }
    rtjx_t* handle();                                                       // creates the C handle on first use

private:
    void check(int rc);
    void print_barcodes_file();                                             // Junction::print_barcodes, .h:99-111
    std::string bam_, ref_;
    uint32_t min_anchor_length_, min_intron_length_, max_intron_length_;
    std::string output_file_, output_barcodes_file_, region_;
    int strandness_;
    std::string strand_tag_, barcode_tag_;
    int device_, n_threads_, shard_rank_, shard_world_;
    rtjx_t* h_;
};

#endif  // RTJX_JUNCTIONS_EXTRACTOR_H_
