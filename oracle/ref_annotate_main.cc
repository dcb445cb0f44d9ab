// oracle/ref_annotate_main.cc — TEST INFRASTRUCTURE ONLY.
// Driver (ours) around the UNMODIFIED reference classes JunctionsAnnotator / GtfParser / BedFile, compiled by
// oracle/Makefile from the sources where they lie under /root/reference.  It reproduces `regtools junctions annotate`
// exactly as the reference's CLI glue does (src/junctions/junctions_main.cc:61-92: parse_options -> load_gtf ->
// open_junctions -> header -> per line adjust_junction_ends, get_splice_site, annotate_junction_with_gtf, print;
// help -> exit 0, runtime_error -> exit 1) without the cmake-generated version.h.
//   regtools_ref_annotate junctions annotate [-S] [-o out] junctions.bed ref.fa annotations.gtf
#include "common.h"
#include "junctions_annotator.h"
#include <cstring>

// The driver loop: every call below is a method of the unmodified reference classes, in the order junctions_main.cc:68-82 uses.
static int annotate_file(JunctionsAnnotator& a) {
    std::ofstream tsv;
    a.load_gtf();
    a.open_junctions();
    a.set_ofstream_object(tsv);
    AnnotatedJunction rec;
    rec.reset();
    rec.print_header(tsv);
    int done = 0;
    for (; a.get_single_junction(rec); ++done) {
        a.adjust_junction_ends(rec);
        a.get_splice_site(rec);
        a.annotate_junction_with_gtf(rec);
        rec.print(tsv);
        rec.reset();
    }
    a.close_ofstream();
    std::cerr << std::endl << "Annotated " << done << " lines." << std::endl;
    a.close_junctions();
    return done;
}

int main(int argc, char** argv) {
    int skip;
    if (argc >= 3 && !strcmp(argv[1], "junctions") && !strcmp(argv[2], "annotate")) skip = 2;
    else if (argc >= 2 && !strcmp(argv[1], "annotate")) skip = 1;
    else { std::cerr << "usage: regtools_ref_annotate junctions annotate [options] junctions.bed ref.fa annotations.gtf\n"; return 1; }
    JunctionsAnnotator annotator;
    int rc = 0;
    try {
        annotator.parse_options(argc - skip, argv + skip);
        annotate_file(annotator);
    } catch (const common::cmdline_help_exception& help) {      // -h: exit 0 (junctions_main.cc:84-86)
        std::cerr << help.what() << std::endl;
    } catch (const std::runtime_error& err) {                   // exit 1 (:87-90)
        std::cerr << err.what() << std::endl;
        rc = 1;
    }
    return rc;
}
