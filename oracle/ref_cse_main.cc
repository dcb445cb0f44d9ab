// oracle/ref_cse_main.cc — TEST INFRASTRUCTURE ONLY (never linked into the product).
// The per-variant junction loop of `cis-splice-effects identify`
// (/root/reference/src/cis-splice-effects/cis_splice_effects_identifier.cc:288-299) around the UNMODIFIED reference classes:
// for every line `region <TAB> cis_effect_start <TAB> cis_effect_end` of the regions file, in order, one JunctionsExtractor
// (8-arg ctor) on that region, get_all_junctions(), the partial-overlap window test (:294-295) and the insert into
// set<Junction> / map<Junction, set<variant>> — whose ordering is the reference's own: Junction has no operator<, so both
// containers compare through the implicit Junction -> AnnotatedJunction conversion (junctions_annotator.h:155-177: chrom,
// start, end; strand-blind; first insert wins).  The VCF / GTF side of that command (which variant gets which window) is not
// on the extract path and is replaced by the regions file.
//   regtools_ref_cse <bam> <strandness> <tag> <min_anchor> <min_intron> <max_intron> <regions.tsv>
// prints, in set order: chrom start end name read_count strand thick_start thick_end <TAB> variant indices (comma separated)
#include "common.h"
#include "junctions_extractor.h"
#include "junctions_annotator.h"
#include <cstdlib>
#include <fstream>
#include <map>
#include <set>
#include <sstream>

int main(int argc, char** argv) {
    if (argc < 8) { std::cerr << "usage: regtools_ref_cse bam strandness tag min_anchor min_intron max_intron regions.tsv\n"; return 2; }
    std::set<Junction> unique_junctions_;
    std::map<Junction, std::set<int> > junction_to_variant_;
    std::ifstream in(argv[7]);
    std::string line;
    int vi = 0;
    try {
        while (std::getline(in, line)) {
            std::stringstream ss(line);
            std::string region; unsigned long cs, ce;
            if (!(ss >> region >> cs >> ce)) continue;
            JunctionsExtractor je1(argv[1], region, atoi(argv[2]), argv[3], (uint32_t)strtoul(argv[4], 0, 10),
                                   (uint32_t)strtoul(argv[5], 0, 10), (uint32_t)strtoul(argv[6], 0, 10), "NA");
            je1.identify_junctions_from_BAM();
            std::vector<Junction> junctions = je1.get_all_junctions();
            for (size_t i = 0; i < junctions.size(); i++) {
                if ((junctions[i].start >= cs && junctions[i].start <= ce) || (junctions[i].end <= ce && junctions[i].end >= cs)) {
                    unique_junctions_.insert(junctions[i]);
                    junction_to_variant_[junctions[i]].insert(vi);
                }
            }
            ++vi;
        }
    } catch (const std::runtime_error& e) {
        std::cerr << e.what() << std::endl;
        return 1;
    }
    for (std::set<Junction>::iterator j1 = unique_junctions_.begin(); j1 != unique_junctions_.end(); j1++) {
        const Junction& j = *j1;
        std::cout << j.chrom << "\t" << j.start << "\t" << j.end << "\t" << j.name << "\t" << j.read_count << "\t" << j.strand << "\t"
                  << j.thick_start << "\t" << j.thick_end << "\t";
        const std::set<int>& vs = junction_to_variant_[j];
        for (std::set<int>::const_iterator v = vs.begin(); v != vs.end(); ++v) std::cout << (v == vs.begin() ? "" : ",") << *v;
        std::cout << "\n";
    }
    return 0;
}
