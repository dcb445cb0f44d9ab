// oracle/ref_main.cc — TEST INFRASTRUCTURE ONLY.
// A 30-line driver (ours) around the UNMODIFIED reference class, compiled by oracle/Makefile
// from /root/reference/src/junctions/junctions_extractor.cc + vendored htslib where they lie.
// It reproduces `regtools junctions extract` exactly as the reference's CLI glue does
// (src/junctions/junctions_main.cc:45-59: parse_options -> identify_junctions_from_BAM ->
// print_all_junctions; help -> exit 0, runtime_error -> exit 1) without dragging in the
// annotate/GTF/cis-ase subsystems or the cmake-generated version.h.
//
// Extra mode used by the parity tests for the second caller of the hot path
// (src/cis-splice-effects/cis_splice_effects_identifier.cc:288-290):
//   regtools_ref ctor <bam> <region> <strandness> <tag> <min_anchor> <min_intron> <max_intron>
// constructs the 8-arg ctor, runs, and dumps get_all_junctions() (unfiltered) as BED12 plus
// two trailing columns (has_left, has_right).
#include "common.h"
#include "junctions_extractor.h"
#include <cstring>
#include <cstdlib>

static int run_extract(int argc, char** argv) {
    JunctionsExtractor extract;
    try {
        extract.parse_options(argc, argv);
        extract.identify_junctions_from_BAM();
        extract.print_all_junctions();
    } catch (const common::cmdline_help_exception& e) {
        std::cerr << e.what() << std::endl;
        return 0;
    } catch (const std::runtime_error& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}

static int run_ctor(int argc, char** argv) {
    if (argc < 9) { std::cerr << "ctor needs 7 args\n"; return 2; }
    try {
        JunctionsExtractor je(argv[2], argv[3], atoi(argv[4]), argv[5],
                              (uint32_t)strtoul(argv[6], 0, 10), (uint32_t)strtoul(argv[7], 0, 10),
                              (uint32_t)strtoul(argv[8], 0, 10), "NA");
        je.identify_junctions_from_BAM();
        std::vector<Junction> v = je.get_all_junctions();
        for (size_t i = 0; i < v.size(); ++i) {
            const Junction& j = v[i];
            std::cout << j.chrom << "\t" << j.thick_start << "\t" << j.thick_end << "\t" << j.name << "\t"
                      << j.read_count << "\t" << j.strand << "\t" << j.start << "\t" << j.end << "\t"
                      << (int)j.has_left_min_anchor << "\t" << (int)j.has_right_min_anchor << "\n";
        }
    } catch (const std::runtime_error& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    return 0;
}

int main(int argc, char** argv) {
    // accepted spellings: `regtools_ref junctions extract ...`, `regtools_ref extract ...`
    if (argc >= 2 && !strcmp(argv[1], "ctor")) return run_ctor(argc, argv);
    int skip = 0;
    if (argc >= 3 && !strcmp(argv[1], "junctions") && !strcmp(argv[2], "extract")) skip = 2;
    else if (argc >= 2 && !strcmp(argv[1], "extract")) skip = 1;
    else { std::cerr << "usage: regtools_ref junctions extract [options] in.bam\n"; return 1; }
    return run_extract(argc - skip, argv + skip);
}
