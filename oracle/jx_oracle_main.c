/* oracle/jx_oracle_main.c — TEST INFRASTRUCTURE ONLY.
 * CLI around the C restatement, same flags as `regtools junctions extract`
 * (junctions_extractor.cc:42-122), used by tests and by bench.py's CPU legs. */
#define _GNU_SOURCE
#include "jx_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int main(int argc, char** argv) {
    uint32_t a = 8, m = 70, M = 500000; int s = -1; const char *out = NULL, *reg = ".", *tag = "XS", *bc_out = NULL;
    int c;
    while ((c = getopt(argc, argv, "a:m:M:o:r:t:s:b:")) != -1) {
        switch (c) {
        case 'a': a = (uint32_t)atoi(optarg); break;
        case 'm': m = (uint32_t)atoi(optarg); break;
        case 'M': M = (uint32_t)atoi(optarg); break;
        case 'o': out = optarg; break;
        case 'r': reg = optarg; break;
        case 't': tag = optarg; break;
        case 'b': bc_out = optarg; break;       /* writes the REPLAY input (see jxo_write_barcode_replay_path), pipe through bc_replay */
        case 's': s = !strcmp(optarg, "XS") ? 0 : !strcmp(optarg, "RF") ? 1 : !strcmp(optarg, "FR") ? 2 :
                      !strcmp(optarg, "intron-motif") ? 3 : -1; break;
        default: return 1;
        }
    }
    if (optind >= argc || s < 0) { fprintf(stderr, "usage: jx_oracle -s XS|RF|FR|intron-motif [-a -m -M -o -r -t] in.bam [ref.fa]\n"); return 1; }
    if (s == 3 && optind + 1 >= argc) { fprintf(stderr, "Strandness mode 'intron-motif' requires a fasta file!\n\n"); return 1; }
    jxo_t* o = jxo_new(a, m, M, s, tag);
    if (optind + 1 < argc && jxo_set_fasta(o, argv[optind + 1])) { fprintf(stderr, "cannot read %s\n", argv[optind + 1]); return 1; }
    if (bc_out) jxo_enable_barcodes(o, NULL);
    const char* err = NULL;
    if (jxo_extract_bam(o, argv[optind], reg, &err)) { fprintf(stderr, "%s", err ? err : "error\n"); return 1; }
    if (out) jxo_write_bed12_path(o, out); else jxo_write_bed12(o, stdout);
    if (bc_out) jxo_write_barcode_replay_path(o, bc_out);
    fprintf(stderr, "reads=%llu junctions=%zu\n", (unsigned long long)jxo_reads_seen(o), jxo_count(o));
    jxo_free(o);
    return 0;
}
