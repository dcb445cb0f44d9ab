// regtools_b200/csrc/regtools_main.cc — the `regtools` CLI front-end of the B200 path.
//
// Same dispatch, banner and exit codes as /root/reference/src/regtools.cc:36-74 and
// src/junctions/junctions_main.cc:45-107.  `junctions extract` (the hot path) and `junctions annotate` (its downstream
// consumer) are built here; the other sub-commands are outside the scope and say so.
#include <cstring>
#include <iostream>
#include <string>

#include "junctions_annotator.h"
#include "junctions_extractor.h"

using namespace std;

static void version() {
    cerr << endl;
    cerr << "Program:\tregtools" << endl;
    cerr << "Version:\t" << 1 << "." << 0 << "." << 0 << endl;
}

static int usage() {
    cerr << "Usage:" << "\t\t" << "regtools <command> [options]" << endl;
    cerr << "Command:\t" << "junctions\t\tTools that operate on feature junctions (e.g. exon-exon junctions from RNA-seq)." << endl;
    cerr << "\t\t" << "cis-ase\t\t\tTools related to allele specific expression in cis." << endl;
    cerr << "\t\t" << "cis-splice-effects\tTools related to splicing effects of variants." << endl;
    cerr << "\t\t" << "variants\t\tTools that operate on variants." << endl;
    cerr << endl;
    return 0;
}

static int junctions_usage(ostream& out = cout) {
    out << "Usage:\t\t" << "regtools junctions <command> [options]" << endl;
    out << "Command:\t" << "extract\t\tIdentify exon-exon junctions from alignments." << endl;
    out << "\t\tannotate\tAnnotate the junctions." << endl;
    out << endl;
    return 0;
}

static int junctions_extract(int argc, char* argv[]) {
    JunctionsExtractor extract;
    if (const char* d = getenv("RTJX_DEVICE")) extract.set_device(atoi(d));
    if (const char* t = getenv("RTJX_THREADS")) extract.set_threads(atoi(t));
    try {
        extract.parse_options(argc, argv);
        extract.identify_junctions_from_BAM();
        extract.print_all_junctions();
    } catch (const common::cmdline_help_exception& e) {
        cerr << e.what() << endl;
        return 0;
    } catch (const runtime_error& error) {
        cerr << error.what() << endl;
        return 1;
    }
    return 0;
}

static int junctions_annotate(int argc, char* argv[]) {       // junctions_main.cc:61-92
    JunctionsAnnotator anno;
    if (const char* d = getenv("RTJX_DEVICE")) anno.set_device(atoi(d));
    try {
        anno.parse_options(argc, argv);
        anno.annotate_all();                                 // prints "Annotated N lines." itself (chatter_fd)
    } catch (const common::cmdline_help_exception& e) {
        cerr << e.what() << endl;
        return 0;
    } catch (const runtime_error& e) {
        // two GtfParser failures are `cerr << text; exit(1)` without a line end (gtf_parser.cc:52-55,202-206), the rest are
        // runtime_errors printed with endl by the driver (junctions_main.cc:87-90)
        const string msg(e.what());
        if (msg.find("\nUnable to open GTF file.") == 0 || msg.find("Undefined strand for exon") == 0) cerr << msg;
        else cerr << msg << endl;
        return 1;
    }
    return 0;
}

static int not_built(const char* what) {
    cerr << "regtools (B200 build): '" << what << "' is outside the junctions-extract hot path and is not built here." << endl;
    return 1;
}

int main(int argc, char* argv[]) {
    version();
    if (argc > 1) {
        string subcmd(argv[1]);
        if (subcmd == "junctions") {
            if (argc > 2) {
                string sub2(argv[2]);
                if (sub2 == "extract") return junctions_extract(argc - 2, argv + 2);
                if (sub2 == "annotate") return junctions_annotate(argc - 2, argv + 2);
            }
            return junctions_usage();
        }
        if (subcmd == "variants" || subcmd == "cis-splice-effects" || subcmd == "cis-ase") return not_built(argv[1]);
    }
    return usage();
}
