/* oracle/ref_index.c — TEST INFRASTRUCTURE ONLY.
 * Builds <bam>.bai with the reference's own vendored htslib (sam_index_build,
 * /root/reference/src/utils/htslib/sam.c), used to cross-check our BAI writer. */
#include <stdio.h>
#include "htslib/sam.h"
int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: ref_index in.bam\n"); return 2; }
    return bam_index_build(argv[1], 0) == 0 ? 0 : 1;
}
