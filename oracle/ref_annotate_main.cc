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

int main(int argc, char** argv) {
    int skip = 0;
    if (argc >= 3 && !strcmp(argv[1], "junctions") && !strcmp(argv[2], "annotate")) skip = 2;
    else if (argc >= 2 && !strcmp(argv[1], "annotate")) skip = 1;
    else { std::cerr << "usage: regtools_ref_annotate junctions annotate [options] junctions.bed ref.fa annotations.gtf\n"; return 1; }
    argc -= skip; argv += skip;
    JunctionsAnnotator anno;
    AnnotatedJunction line;
    line.reset();
    int linec = 0;
    ofstream out;
    try {
        anno.parse_options(argc, argv);
        anno.load_gtf();
        anno.open_junctions();
        anno.set_ofstream_object(out);
        line.print_header(out);
        while (anno.get_single_junction(line)) {
            anno.adjust_junction_ends(line);
            anno.get_splice_site(line);
            anno.annotate_junction_with_gtf(line);
            line.print(out);
            line.reset();
            linec++;
        }
        anno.close_ofstream();
        cerr << endl << "Annotated " << linec << " lines." << endl;
        anno.close_junctions();
    } catch (const common::cmdline_help_exception& e) {
        cerr << e.what() << endl;
        return 0;
    } catch (const runtime_error& e) {
        cerr << e.what() << endl;
        return 1;
    }
    return 0;
}
