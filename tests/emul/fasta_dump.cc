// tests/emul/fasta_dump.cc — TEST INFRASTRUCTURE ONLY: prints what regtools_b200/csrc/fasta.cc (load_fasta) makes of a FASTA,
// one line per sequence: name <TAB> length <TAB> FNV-1a of the bases <TAB> first 40 bases.  Compared with a plain Python
// reading of the same file in tests/test_host_logic.py.
#include <cstdio>
#include "../../regtools_b200/csrc/fasta.h"
int main(int argc, char** argv) {
    if (argc != 2) return 2;
    rtjx::FastaGenome g; std::string err;
    if (!rtjx::load_fasta(argv[1], &g, &err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    for (size_t i = 0; i < g.names.size(); ++i) {
        unsigned long long h = 1469598103934665603ull;
        for (uint64_t k = 0; k < g.length[i]; ++k) { h ^= g.bases[g.offset[i] + k]; h *= 1099511628211ull; }
        printf("%s\t%llu\t%016llx\t%.*s\n", g.names[i].c_str(), (unsigned long long)g.length[i], h,
               (int)(g.length[i] < 40 ? g.length[i] : 40), (const char*)g.bases.data() + g.offset[i]);
    }
    return 0;
}
