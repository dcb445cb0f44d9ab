// regtools_b200/csrc/junctions_annotator.h — C++ surface of `regtools junctions annotate` over the C ABI.
//
// Mirrors what the reference's CLI glue uses of class JunctionsAnnotator
// (/root/reference/src/junctions/junctions_annotator.h:164-256, driver src/junctions/junctions_main.cc:61-92):
// parse_options / usage with the same flags, texts and exceptions; the per-line loop of the driver (adjust ends, splice
// site, GTF overlap, print) is one call, annotate_all(), because it runs batched on the device (rtjx_annotate).
#ifndef RTJX_JUNCTIONS_ANNOTATOR_H_
#define RTJX_JUNCTIONS_ANNOTATOR_H_
#include <iostream>
#include <string>

class JunctionsAnnotator {
public:
    JunctionsAnnotator() : ref_("NA"), skip_single_exon_genes_(true), output_file_("NA"), device_(0) {}   // .h:198-202
    int parse_options(int argc, char* argv[]);                              // junctions_annotator.cc:405-447
    int usage(std::ostream& out = std::cerr);                               // :450-456
    std::string gtf_file() { return gtf_; }                                 // :400-402
    // load_gtf + open_junctions + set_ofstream_object + print_header + the while loop of junctions_main.cc:68-82;
    // returns the number of annotated lines, throws std::runtime_error with the reference's texts
    int annotate_all();
    void set_device(int device) { device_ = device; }
private:
    std::string junctions_, ref_, gtf_;
    bool skip_single_exon_genes_;
    std::string output_file_;
    int device_;
};

#endif  // RTJX_JUNCTIONS_ANNOTATOR_H_
