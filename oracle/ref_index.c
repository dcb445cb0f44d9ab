/* oracle/ref_index.c — TEST INFRASTRUCTURE ONLY.
 * Builds an index with the reference's own vendored htslib (sam_index_build,
 * /root/reference/src/utils/htslib/sam.c): <bam>.bai (default, used to cross-check our BAI writer) or, with a
 * min_shift argument > 0, <bam>.csi (fixtures for the CSI reader: tests/golden/csi). */
#include <stdio.h>
#include <stdlib.h>
#include "htslib/sam.h"
int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_index in.bam [min_shift]\n"); return 2; }
    return bam_index_build(argv[1], argc > 2 ? atoi(argv[2]) : 0) == 0 ? 0 : 1;
}
