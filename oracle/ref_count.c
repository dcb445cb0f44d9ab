/* oracle/ref_count.c — TEST / BENCH INFRASTRUCTURE ONLY (never linked into the product).
 * Counts the alignments the reference's extractor sees for a region, with the reference's own vendored htslib and the very
 * calls JunctionsExtractor::identify_junctions_from_BAM makes (/root/reference/src/junctions/junctions_extractor.cc:500-535:
 * sam_open, sam_index_load, sam_hdr_read, sam_itr_querys, sam_itr_next).  bench.py's reference arm uses it for the numerator
 * of reads/s so that nothing of the product is loaded there.
 *     ref_count in.bam [region]        prints the count; region defaults to "." (the whole file, as the CLI does) */
#include <stdio.h>
#include <stdlib.h>
#include "htslib/sam.h"
int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_count in.bam [region]\n"); return 2; }
    const char* region = argc > 2 ? argv[2] : ".";
    samFile* in = sam_open(argv[1], "r");
    if (!in) { fprintf(stderr, "Unable to open BAM/SAM file.\n"); return 1; }
    hts_idx_t* idx = sam_index_load(in, argv[1]);
    if (!idx) { fprintf(stderr, "Unable to open BAM/SAM index.\n"); return 1; }
    bam_hdr_t* hdr = sam_hdr_read(in);
    hts_itr_t* it = sam_itr_querys(idx, hdr, region);
    if (!hdr || !it) { fprintf(stderr, "Unable to iterate to region within BAM.\n"); return 1; }
    bam1_t* b = bam_init1();
    unsigned long long n = 0, multi = 0;
    while (sam_itr_next(in, it, b) >= 0) { ++n; multi += b->core.n_cigar > 1; }
    printf("%llu\t%llu\n", n, multi);
    hts_itr_destroy(it); bam_destroy1(b); bam_hdr_destroy(hdr); hts_idx_destroy(idx); sam_close(in);
    return 0;
}
