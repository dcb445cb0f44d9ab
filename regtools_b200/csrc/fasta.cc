// regtools_b200/csrc/fasta.cc — see fasta.h.
#include "fasta.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <thread>
#include <unordered_map>
#include <vector>

namespace rtjx {

int FastaGenome::find(const std::string& name) const {
    for (size_t i = 0; i < names.size(); ++i) if (names[i] == name) return (int)i;
    return -1;
}

namespace {
inline bool is_space(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }
inline bool is_graph(unsigned char c) { return c > 32 && c < 127; }
}  // namespace

// The file is parsed by several threads: byte ranges cut at line starts are scanned in parallel (header lines found, isgraph()
// bytes counted per segment between headers), one sequential pass over the few header events decides names, duplicates and
// output offsets, and the ranges are then copied in parallel.  Result identical to a single sequential pass.
namespace {
struct Segment { size_t begin, end; uint64_t n_bases; uint64_t out_off; bool keep; };   // lines [begin, end) of one sequence inside a chunk
struct HeaderEvent { size_t seg; std::string name; };                                     // segment `seg` of the chunk starts a new sequence
struct Chunk { size_t a, b; std::vector<Segment> segs; std::vector<HeaderEvent> headers; };

inline size_t line_end(const unsigned char* p, size_t j, size_t sz) {
    const void* e = memchr(p + j, '\n', sz - j);
    return e ? (size_t)((const unsigned char*)e - p) : sz;
}
inline uint64_t count_graph(const unsigned char* p, size_t j, size_t le) {
    uint64_t bad = 0;
    for (size_t q = j; q < le; ++q) bad += (uint64_t)((unsigned char)(p[q] - 33) >= 94);
    return (uint64_t)(le - j) - bad;
}
void scan_chunk(const unsigned char* p, size_t sz, Chunk* c) {
    size_t j = c->a;
    c->segs.push_back(Segment{j, j, 0, 0, false});            // continuation of the sequence open at the chunk's start
    while (j < c->b) {
        const size_t le = line_end(p, j, sz);
        if (p[j] == '>') {                                     // fai_build_core starts a record at a '>' at the beginning of a line (faidx.c:103)
            size_t q = j + 1;
            while (q < le && is_space(p[q])) ++q;
            const size_t n0 = q;
            while (q < le && !is_space(p[q])) ++q;
            c->segs.back().end = j;
            c->headers.push_back(HeaderEvent{c->segs.size(), std::string((const char*)p + n0, q - n0)});
            c->segs.push_back(Segment{le < sz ? le + 1 : sz, le < sz ? le + 1 : sz, 0, 0, false});
        } else {
            c->segs.back().n_bases += count_graph(p, j, le);
        }
        j = le < sz ? le + 1 : sz;
        c->segs.back().end = j;
    }
}
void copy_chunk(const unsigned char* p, size_t sz, const Chunk& c, uint8_t* out) {
    for (const Segment& s : c.segs) {
        if (!s.keep || !s.n_bases) continue;
        uint8_t* o = out + s.out_off;
        for (size_t j = s.begin; j < s.end;) {
            const size_t le = std::min(line_end(p, j, sz), s.end);
            if (count_graph(p, j, le) == le - j) { memcpy(o, p + j, le - j); o += le - j; }
            else for (size_t q = j; q < le; ++q) if (is_graph(p[q])) *o++ = p[q];
            j = le + 1;
        }
    }
}
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
    g->names.clear(); g->offset.clear(); g->length.clear(); g->bases.clear();

    // ---- chunks cut at line starts
    unsigned hw = std::thread::hardware_concurrency();
    const size_t n_chunks = std::max<size_t>(1, std::min<size_t>({(size_t)(hw ? hw : 4), (size_t)16, sz / (4u << 20) + 1}));
    std::vector<Chunk> chunks(n_chunks);
    size_t prev = 0;
    for (size_t k = 0; k < n_chunks; ++k) {
        size_t b = k + 1 == n_chunks ? sz : std::max(prev, sz / n_chunks * (k + 1));
        if (b < sz) { const size_t le = line_end(p, b, sz); b = le < sz ? le + 1 : sz; }
        chunks[k].a = prev; chunks[k].b = b;
        prev = b;
    }
    auto parallel = [&](auto fn) {
        std::vector<std::thread> th;
        for (size_t k = 1; k < n_chunks; ++k) th.emplace_back(fn, k);
        fn((size_t)0);
        for (std::thread& t : th) t.join();
    };
    parallel([&](size_t k) { scan_chunk(p, sz, &chunks[k]); });

    // ---- names, duplicates (the first sequence of a name wins, faidx.c:118-124) and output offsets
    std::unordered_map<std::string, int> seen;
    uint64_t total = 0;
    bool keep = false;                                         // bytes before the first header belong to no sequence
    int cur = -1;
    for (Chunk& c : chunks) {
        size_t h = 0;
        for (size_t si = 0; si < c.segs.size(); ++si) {
            if (h < c.headers.size() && c.headers[h].seg == si) {
                if (cur >= 0) g->length[(size_t)cur] = total - g->offset[(size_t)cur];
                const std::string& name = c.headers[h].name;
                keep = seen.count(name) == 0;
                cur = -1;
                if (keep) {
                    seen[name] = (int)g->names.size();
                    cur = (int)g->names.size();
                    g->names.push_back(name); g->offset.push_back(total); g->length.push_back(0);
                }
                ++h;
            }
            Segment& s = c.segs[si];
            s.keep = keep; s.out_off = total;
            if (keep) total += s.n_bases;
        }
    }
    if (cur >= 0) g->length[(size_t)cur] = total - g->offset[(size_t)cur];
    if (g->names.empty()) { munmap((void*)p, sz); *err = "no sequence in FASTA " + path; return false; }
    g->bases.resize((size_t)total + 16);
    uint8_t* out = g->bases.data();
    parallel([&](size_t k) { copy_chunk(p, sz, chunks[k], out); });
    memset(out + total, 0, 16);
    munmap((void*)p, sz);
    return true;
}

}  // namespace rtjx
