// regtools_b200/csrc/junctions_annotator.cc — see junctions_annotator.h.
#include "junctions_annotator.h"

#include <unistd.h>

#include <sstream>
#include <stdexcept>

#include "../../include/rtjx.h"
#include "junctions_extractor.h"      // common::cmdline_help_exception

using namespace std;

int JunctionsAnnotator::usage(ostream& out) {
    out << "Usage:\t\t" << "regtools junctions annotate [options] junctions.bed ref.fa annotations.gtf" << endl;
    out << "Options:\t" << "-S include single exon genes" << endl;
    out << "\t\t" << "-o FILE\tThe file to write output to. [STDOUT]" << endl;
    out << endl;
    return 0;
}

int JunctionsAnnotator::parse_options(int argc, char* argv[]) {
    optind = 1;
    int c;
    stringstream help_ss;
    while ((c = getopt(argc, argv, "So:h")) != -1) {
        switch (c) {
        case 'S': skip_single_exon_genes_ = false; break;
        case 'o': output_file_ = string(optarg); break;
        case 'h': usage(help_ss); throw common::cmdline_help_exception(help_ss.str());
        default: usage(); throw runtime_error("Error parsing inputs!(1)\n\n");
        }
    }
    if (argc - optind >= 3) {
        junctions_ = string(argv[optind++]);
        ref_ = string(argv[optind++]);
        gtf_ = string(argv[optind++]);
    }
    if (optind < argc || ref_ == "NA" || junctions_.empty() || gtf_.empty()) {
        usage();
        throw runtime_error("Error parsing inputs!(2)\n\n");
    }
    cerr << "Reference: " << ref_ << endl;
    cerr << "GTF: " << gtf_ << endl;
    cerr << "Junctions: " << junctions_ << endl;
    if (skip_single_exon_genes_) cerr << "Skipping single exon genes." << endl;
    if (output_file_ != "NA") cerr << "Output file: " << output_file_ << endl;
    cerr << endl;
    return 0;
}

int JunctionsAnnotator::annotate_all() {
    rtjx_annotate_params p;
    rtjx_annotate_params_default(&p);
    p.junctions_bed = junctions_.c_str(); p.fasta = ref_.c_str(); p.gtf = gtf_.c_str();
    p.include_single_exon = skip_single_exon_genes_ ? 0 : 1;
    p.device = device_;
    p.chatter_fd = STDERR_FILENO;
    int fd = STDOUT_FILENO;
    if (output_file_ != "NA") { fd = -1; p.out_path = output_file_.c_str(); }   // opened by the library after the GTF is in (:68-71)
    else cout.flush();
    char err[512];
    uint64_t n = 0;
    cerr.flush();
    const int rc = rtjx_annotate(&p, fd, &n, err, sizeof err);
    if (rc != RTJX_OK) throw runtime_error(err[0] ? string(err) : string(rtjx_strerror(rc)));
    return (int)n;
}
