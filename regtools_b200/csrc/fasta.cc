// regtools_b200/csrc/fasta.cc — see fasta.h.
#include "fasta.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstring>
#include <unordered_map>

namespace rtjx {

int FastaGenome::find(const std::string& name) const {
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) return (int)i;
    return -1;
}

namespace {
inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
inline bool is_graph(unsigned char c) { return c > 32 && c < 127; }
}  // namespace

bool load_fasta(const std::string& path, FastaGenome* g, std::string* err) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { *err = "cannot open FASTA " + path; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { ::close(fd); *err = "empty or unreadable FASTA " + path; return false; }
    const size_t sz = (size_t)st.st_size;
    const unsigned char* p = (const unsigned char*)mmap(nullptr, sz, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) { *err = "cannot map FASTA " + path; return false; }
    if (sz >= 2 && p[0] == 0x1f && p[1] == 0x8b) { munmap((void*)p, sz); *err = "compressed FASTA is not supported by the B200 path"; return false; }
    madvise((void*)p, sz, MADV_SEQUENTIAL);
    g->names.clear(); g->offset.clear(); g->length.clear(); g->bases.clear();
    g->bases.reserve(sz + 16);
    std::unordered_map<std::string, int> seen;
    size_t i = 0;
    // fai_build_core starts a record at a '>' met at the beginning of a line (faidx.c:103)
    while (i < sz) {
        if (p[i] != '>') { const void* nl = memchr(p + i, '\n', sz - i); i = nl ? (size_t)((const unsigned char*)nl - p) + 1 : sz; continue; }
        size_t j = i + 1;
        while (j < sz && p[j] != '\n' && is_space(p[j])) ++j;
        const size_t n0 = j;
        while (j < sz && !is_space(p[j])) ++j;
        std::string name((const char*)p + n0, j - n0);
        const void* nl = memchr(p + j, '\n', sz - j);
        j = nl ? (size_t)((const unsigned char*)nl - p) + 1 : sz;
        const bool dup = seen.count(name) != 0;
        const uint64_t off = g->bases.size();
        // sequence lines up to the next header line
        while (j < sz && p[j] != '>') {
            const void* e = memchr(p + j, '\n', sz - j);
            const size_t le = e ? (size_t)((const unsigned char*)e - p) : sz;
            if (!dup) {
                // a sequence line is normally all isgraph(): test the line in one vectorisable pass, then copy it whole
                unsigned bad = 0;
                for (size_t q = j; q < le; ++q) bad |= (unsigned)((unsigned char)(p[q] - 33) >= 94);
                if (!bad) g->bases.insert(g->bases.end(), p + j, p + le);
                else for (size_t q = j; q < le; ++q) if (is_graph(p[q])) g->bases.push_back(p[q]);
            }
            j = e ? le + 1 : sz;
        }
        if (!dup) {
            seen[name] = (int)g->names.size();
            g->names.push_back(name); g->offset.push_back(off); g->length.push_back(g->bases.size() - off);
        }
        i = j;
    }
    munmap((void*)p, sz);
    if (g->names.empty()) { *err = "no sequence in FASTA " + path; return false; }
    g->bases.resize(g->bases.size() + 16, 0);
    return true;
}

}  // namespace rtjx
