// tools/bamgen.cc — bench/test infrastructure: deterministic synthetic RNA-seq BAM generator with
// multi-threaded BGZF deflate, and a BAI indexer.  Not part of the product path.
//
//   bamgen gen --out x.bam --config c2|c2xN|c3|tiny --reads N [--seed 1234] [--level 6] [--threads T] [--qual8] [--barcodes N]
//   bamgen index x.bam            (writes x.bam.bai)
//
// Workload shape follows SURVEY.md §8(d): junction catalog of one intron per 12 kb with Zipf(1)
// expression weights (hot junctions get 1e5-1e6 reads), log-uniform intron lengths with 0.5 %
// outside [70, 500000], 8 % / 12 % spliced reads, 10 % of those with a second N, S/D/I/X/=/H mixes
// on unspliced reads, MAPQ and flag mixes, XS:A on 95 % of spliced reads, random SEQ/QUAL.
// The output is independent of the thread count (work is cut into fixed windows, each seeded from
// its id).  The index has the BAI layout htslib writes (bins + 16 kb linear index + pseudo-bin
// 37450; semantics of /root/reference/src/utils/htslib/hts.c:1288-1350 hts_idx_push).
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Rng {
    uint64_t s[2];
    explicit Rng(uint64_t seed) {
        auto sm = [&seed]() {
            uint64_t z = (seed += 0x9E3779B97F4A7C15ull);
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            return z ^ (z >> 31);
        };
        s[0] = sm(); s[1] = sm();
    }
    inline uint64_t next() {   // xoroshiro128+
        uint64_t a = s[0], b = s[1], r = a + b;
        b ^= a; s[0] = ((a << 24) | (a >> 40)) ^ b ^ (b << 16); s[1] = (b << 37) | (b >> 27);
        return r;
    }
    inline double uni() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    inline uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
};

// ------------------------------------------------------------------------------------------ BAI
struct Chunk { uint64_t u, v; };
constexpr uint32_t META_BIN = 37450;

inline int reg2bin(int64_t beg, int64_t end) {
    --end;
    if (beg >> 14 == end >> 14) return (int)(((1 << 15) - 1) / 7 + (beg >> 14));
    if (beg >> 17 == end >> 17) return (int)(((1 << 12) - 1) / 7 + (beg >> 17));
    if (beg >> 20 == end >> 20) return (int)(((1 << 9) - 1) / 7 + (beg >> 20));
    if (beg >> 23 == end >> 23) return (int)(((1 << 6) - 1) / 7 + (beg >> 23));
    if (beg >> 26 == end >> 26) return (int)(((1 << 3) - 1) / 7 + (beg >> 26));
    return 0;
}

class BaiBuilder {
public:
    explicit BaiBuilder(int n_ref) : refs_((size_t)n_ref) {}
    void set_start(uint64_t voff) { last_off_ = off_beg_ = voff; }
    // one record, in file order; `off_after` = virtual offset right after the record
    void push(int tid, int64_t beg, int64_t end, uint64_t off_after, bool mapped) {
        if (tid < 0) { beg = -1; end = 0; }
        if (last_tid_ != tid) { last_tid_ = tid; last_bin_ = 0xffffffffu; }
        if (tid >= 0) {
            if (mapped) {
                std::vector<uint64_t>& l = refs_[(size_t)tid].lin;
                size_t b = (size_t)(beg >> 14), e = (size_t)((end - 1) >> 14);
                if (e < b) e = b;
                if (l.size() < e + 1) l.resize(e + 1, UINT64_MAX);
                for (size_t i = b; i <= e; ++i) if (l[i] == UINT64_MAX) l[i] = last_off_;
            }
        } else ++n_no_coor_;
        uint32_t bin = (uint32_t)reg2bin(beg, end);
        if (last_bin_ != bin) {
            if (save_bin_ != 0xffffffffu) add_chunk(save_tid_, save_bin_, save_off_, last_off_);
            if (last_bin_ == 0xffffffffu && save_bin_ != 0xffffffffu) {       // contig changed: meta bin
                off_end_ = last_off_;
                add_chunk(save_tid_, META_BIN, off_beg_, off_end_);
                add_chunk(save_tid_, META_BIN, n_mapped_, n_unmapped_);
                n_mapped_ = n_unmapped_ = 0; off_beg_ = off_end_;
            }
            save_off_ = last_off_; save_bin_ = last_bin_ = bin; save_tid_ = tid;
        }
        if (mapped) ++n_mapped_; else ++n_unmapped_;
        last_off_ = off_after;
    }
    void finish(uint64_t final_off) {
        if (save_tid_ >= 0 && save_bin_ != 0xffffffffu) {
            add_chunk(save_tid_, save_bin_, save_off_, final_off);
            add_chunk(save_tid_, META_BIN, off_beg_, final_off);
            add_chunk(save_tid_, META_BIN, n_mapped_, n_unmapped_);
        }
        for (auto& r : refs_) {
            uint64_t off0 = 0;
            auto m = r.bins.find(META_BIN);
            if (m != r.bins.end()) off0 = m->second[0].u;
            size_t l = 0;
            for (; l < r.lin.size() && r.lin[l] == UINT64_MAX; ++l) r.lin[l] = off0;
            for (; l < r.lin.size(); ++l) if (r.lin[l] == UINT64_MAX) r.lin[l] = r.lin[l - 1];
            for (auto& kv : r.bins) {                       // merge chunks touching inside one BGZF block
                if (kv.first >= 37449u) continue;
                std::vector<Chunk>& c = kv.second;
                size_t mm = 0;
                for (size_t i = 1; i < c.size(); ++i) {
                    if (c[mm].v >> 16 >= c[i].u >> 16) { if (c[mm].v < c[i].v) c[mm].v = c[i].v; }
                    else c[++mm] = c[i];
                }
                c.resize(mm + 1);
            }
        }
    }
    bool save(const std::string& path) const {
        FILE* f = fopen(path.c_str(), "wb");
        if (!f) return false;
        fwrite("BAI\1", 1, 4, f);
        int32_t n = (int32_t)refs_.size();
        fwrite(&n, 4, 1, f);
        for (const auto& r : refs_) {
            int32_t nb = (int32_t)r.bins.size();
            fwrite(&nb, 4, 1, f);
            for (const auto& kv : r.bins) {
                uint32_t bin = kv.first; int32_t nc = (int32_t)kv.second.size();
                fwrite(&bin, 4, 1, f); fwrite(&nc, 4, 1, f);
                fwrite(kv.second.data(), 16, kv.second.size(), f);
            }
            int32_t ni = (int32_t)r.lin.size();
            fwrite(&ni, 4, 1, f);
            fwrite(r.lin.data(), 8, r.lin.size(), f);
        }
        fwrite(&n_no_coor_, 8, 1, f);
        return fclose(f) == 0;
    }
private:
    struct Ref { std::map<uint32_t, std::vector<Chunk>> bins; std::vector<uint64_t> lin; };
    void add_chunk(int tid, uint32_t bin, uint64_t u, uint64_t v) {
        if (tid < 0) return;
        refs_[(size_t)tid].bins[bin].push_back(Chunk{u, v});
    }
    std::vector<Ref> refs_;
    int last_tid_ = -2, save_tid_ = -2;
    uint32_t last_bin_ = 0xffffffffu, save_bin_ = 0xffffffffu;
    uint64_t save_off_ = 0, last_off_ = 0, off_beg_ = 0, off_end_ = 0, n_mapped_ = 0, n_unmapped_ = 0, n_no_coor_ = 0;
};

// ------------------------------------------------------------------------------------------ BGZF
// Compresses `n` bytes into one BGZF block appended to `out`.
void bgzf_block(z_stream& zs, const uint8_t* src, uint32_t n, int level, std::vector<uint8_t>& out) {
    (void)level;
    size_t o = out.size();
    out.resize(o + 18 + compressBound(n) + 8 + 64);
    uint8_t* h = out.data() + o;
    static const uint8_t hdr[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
    memcpy(h, hdr, 16);
    deflateReset(&zs);
    zs.next_in = const_cast<Bytef*>(src); zs.avail_in = n;
    zs.next_out = h + 18; zs.avail_out = (uInt)(out.size() - o - 18 - 8);
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "deflate failed\n"); exit(2); }
    uint32_t clen = (uint32_t)zs.total_out;
    uint32_t bsize = 18 + clen + 8;
    if (bsize > 0x10000) { fprintf(stderr, "BGZF block too large\n"); exit(2); }
    uint16_t bs = (uint16_t)(bsize - 1);
    memcpy(h + 16, &bs, 2);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), src, n);
    memcpy(h + 18 + clen, &crc, 4); memcpy(h + 18 + clen + 4, &n, 4);
    out.resize(o + bsize);
}

constexpr uint32_t BLOCK_IN = 0xff00;
const uint8_t EOF_BLOCK[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};

// ------------------------------------------------------------------------------------------ model
struct Contig { std::string name; int64_t len; };
struct Junc {
    int32_t tid; int64_t donor; uint32_t len; char strand; uint32_t exon2; uint32_t len2; uint64_t n_reads; uint64_t id;
};
struct Config {
    std::vector<Contig> contigs; uint32_t read_len; double spliced; bool paired_only; int64_t window;
};

Config make_config(const std::string& name) {
    Config c;
    c.window = 1 << 20;
    if (name == "c2") {
        c.contigs = {{"chr1", 248956422}};
        c.read_len = 101; c.spliced = 0.08; c.paired_only = false;
    } else if (name.rfind("c2x", 0) == 0 && atoi(name.c_str() + 3) > 0) {
        // weak-scaling shape: N contigs, each the c2 chromosome (one contig per GPU when sharded N ways)
        const int n = atoi(name.c_str() + 3);
        for (int i = 0; i < n; ++i) c.contigs.push_back({"chr" + std::to_string(i + 1), 248956422});
        c.read_len = 101; c.spliced = 0.08; c.paired_only = false;
    } else if (name == "c3") {
        static const int64_t L[24] = {248956422, 242193529, 198295559, 190214555, 181538259, 170805979, 159345973, 145138636,
                                      138394717, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345,
                                      83257441, 80373285, 58617616, 64444167, 46709983, 50818468, 156040895, 57227415};
        for (int i = 0; i < 22; ++i) c.contigs.push_back({"chr" + std::to_string(i + 1), L[i]});
        c.contigs.push_back({"chrX", L[22]}); c.contigs.push_back({"chrY", L[23]});
        c.read_len = 150; c.spliced = 0.12; c.paired_only = true;
    } else if (name == "tiny") {      // three small contigs whose lexicographic order differs from tid order
        c.contigs = {{"1", 3000000}, {"10", 2000000}, {"2", 2500000}};
        c.read_len = 76; c.spliced = 0.15; c.paired_only = false; c.window = 1 << 18;
    } else { fprintf(stderr, "unknown config %s\n", name.c_str()); exit(2); }
    return c;
}

struct Rec { int32_t pos; uint32_t order; uint16_t flag; uint8_t mapq; uint8_t xs; uint8_t n_ops; uint32_t ops[8]; };

struct Task {
    int32_t tid; int64_t ws, we; uint64_t n_unspliced; size_t j_lo, j_hi; uint64_t id;
    // outputs
    std::vector<uint8_t> comp;                         // BGZF blocks
    struct IdxRec { int32_t beg, end; uint64_t voff_after; };   // relative virtual offsets
    std::vector<IdxRec> idx;
    bool done = false;
};

struct Gen {
    Config cfg; uint64_t seed; int level; bool qual8; uint64_t n_barcodes = 0;
    std::vector<Junc> cat;                             // sorted by (tid, donor)

    void junction_reads(const Junc& j, int64_t ws, int64_t we, std::vector<Rec>& out) const {
        Rng r(seed * 0x100000001B3ull + j.id * 0x9E3779B97F4A7C15ull + 17);
        const uint32_t L = cfg.read_len;
        for (uint64_t k = 0; k < j.n_reads; ++k) {
            Rec rec; memset(&rec, 0, sizeof rec);
            uint32_t n = 0;
            double u = r.uni();
            uint32_t lead_s = 0, trail_s = 0;
            if (u < 0.03) lead_s = 1 + r.below(8); else if (u < 0.05) trail_s = 1 + r.below(8);
            uint32_t body = L - lead_s - trail_s;
            bool two = r.uni() < 0.10 && body > j.exon2 + 2;
            uint32_t a, rest;
            if (two) { a = 1 + r.below(body - j.exon2 - 1); rest = body - a - j.exon2; }
            else { a = 1 + r.below(body - 1); rest = body - a; }
            if (lead_s) rec.ops[n++] = lead_s << 4 | 4;
            rec.ops[n++] = a << 4 | 0;
            rec.ops[n++] = j.len << 4 | 3;
            if (two) { rec.ops[n++] = j.exon2 << 4 | 0; rec.ops[n++] = j.len2 << 4 | 3; }
            rec.ops[n++] = rest << 4 | 0;
            if (trail_s) rec.ops[n++] = trail_s << 4 | 4;
            rec.n_ops = (uint8_t)n;
            int64_t pos = j.donor - a;
            flags_mapq(r, rec);
            rec.xs = r.uni() < 0.95 ? (uint8_t)j.strand : 0;
            rec.order = (uint32_t)k;
            if (pos >= ws && pos < we) { rec.pos = (int32_t)pos; out.push_back(rec); }
        }
    }
    void flags_mapq(Rng& r, Rec& rec) const {
        static const uint16_t paired[4] = {99, 147, 83, 163};
        static const uint16_t single[2] = {0, 16};
        uint16_t f = (cfg.paired_only || r.uni() < 0.7) ? paired[r.below(4)] : single[r.below(2)];
        double u = r.uni();
        if (u < 0.01) f |= 256; else if (u < 0.02) f |= 1024;
        rec.flag = f;
        double q = r.uni();
        rec.mapq = q < 0.5 ? 60 : q < 0.7 ? 255 : q < 0.8 ? 0 : q < 0.9 ? 1 : 3;
    }
    void unspliced_reads(const Task& t, std::vector<Rec>& out) const {
        Rng r(seed * 0x100000001B3ull + t.id * 0xD6E8FEB86659FD93ull + 5);
        const uint32_t L = cfg.read_len;
        int64_t hi = std::min<int64_t>(t.we, cfg.contigs[(size_t)t.tid].len - 2 * (int64_t)L);
        if (hi <= t.ws) hi = t.ws + 1;
        for (uint64_t k = 0; k < t.n_unspliced; ++k) {
            Rec rec; memset(&rec, 0, sizeof rec);
            rec.pos = (int32_t)(t.ws + (int64_t)(r.uni() * (double)(hi - t.ws)));
            uint32_t n = 0; double u = r.uni();
            if (u < 0.025) { uint32_t s = 1 + r.below(20); rec.ops[n++] = s << 4 | 4; rec.ops[n++] = (L - s) << 4 | 0; }
            else if (u < 0.05) { uint32_t s = 1 + r.below(20); rec.ops[n++] = (L - s) << 4 | 0; rec.ops[n++] = s << 4 | 4; }
            else if (u < 0.07) { uint32_t a = 10 + r.below(L - 20); rec.ops[n++] = a << 4 | 0; rec.ops[n++] = (1 + r.below(5)) << 4 | 2; rec.ops[n++] = (L - a) << 4 | 0; }
            else if (u < 0.09) { uint32_t a = 10 + r.below(L - 25), i = 1 + r.below(4); rec.ops[n++] = a << 4 | 0; rec.ops[n++] = i << 4 | 1; rec.ops[n++] = (L - a - i) << 4 | 0; }
            else if (u < 0.095) { uint32_t a = 10 + r.below(L - 20); rec.ops[n++] = a << 4 | 7; rec.ops[n++] = 1 << 4 | 8; rec.ops[n++] = (L - a - 1) << 4 | 7; }
            else if (u < 0.097) { rec.ops[n++] = 10 << 4 | 5; rec.ops[n++] = L << 4 | 0; }
            else rec.ops[n++] = L << 4 | 0;
            rec.n_ops = (uint8_t)n;
            flags_mapq(r, rec);
            rec.order = (uint32_t)k | 0x80000000u;
            out.push_back(rec);
        }
    }
    void run_task(Task& t) const {
        std::vector<Rec> recs;
        recs.reserve((size_t)t.n_unspliced + 1024);
        unspliced_reads(t, recs);
        for (size_t j = t.j_lo; j < t.j_hi; ++j) junction_reads(cat[j], t.ws, t.we, recs);
        std::stable_sort(recs.begin(), recs.end(), [](const Rec& a, const Rec& b) { return a.pos < b.pos; });
        // serialise + compress
        Rng r(seed * 0x100000001B3ull + t.id * 0xA24BAED4963EE407ull + 99);
        const uint32_t L = cfg.read_len;
        std::vector<uint8_t> raw; raw.reserve(BLOCK_IN + 1024);
        z_stream zs; memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        t.idx.reserve(recs.size());
        uint64_t coff_rel = 0;
        auto flush_block = [&](uint32_t n) {
            size_t before = t.comp.size();
            bgzf_block(zs, raw.data(), n, level, t.comp);
            coff_rel += t.comp.size() - before;
            raw.erase(raw.begin(), raw.begin() + n);
        };
        uint8_t rec_buf[512];
        uint64_t serial = 0;
        for (const Rec& rc : recs) {
            uint8_t* p = rec_buf + 4;
            int32_t end = rc.pos;
            for (uint32_t k = 0; k < rc.n_ops; ++k) { uint32_t op = rc.ops[k] & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += (int32_t)(rc.ops[k] >> 4); }
            char name[24]; int l_name = snprintf(name, sizeof name, "w%05u.%07llu", (unsigned)(t.id % 100000), (unsigned long long)serial++) + 1;
            int32_t tid = t.tid, pos = rc.pos, l_seq = (int32_t)L, mtid = (rc.flag & 1) ? t.tid : -1;
            int32_t mpos = (rc.flag & 1) ? std::max(0, rc.pos + (int32_t)r.below(400) - 200) : -1, tlen = (rc.flag & 1) ? (int32_t)r.below(600) - 300 : 0;
            uint32_t bin_mq_nl = (uint32_t)reg2bin(rc.pos, end) << 16 | (uint32_t)rc.mapq << 8 | (uint32_t)l_name;
            uint32_t flag_nc = (uint32_t)rc.flag << 16 | rc.n_ops;
            memcpy(p, &tid, 4); memcpy(p + 4, &pos, 4); memcpy(p + 8, &bin_mq_nl, 4); memcpy(p + 12, &flag_nc, 4);
            memcpy(p + 16, &l_seq, 4); memcpy(p + 20, &mtid, 4); memcpy(p + 24, &mpos, 4); memcpy(p + 28, &tlen, 4);
            p += 32;
            memcpy(p, name, (size_t)l_name); p += l_name;
            memcpy(p, rc.ops, 4u * rc.n_ops); p += 4u * rc.n_ops;
            for (uint32_t i = 0; i < (L + 1) / 2; i += 8) {       // random 4-bit bases from {1,2,4,8}
                uint64_t x = r.next(), y = 0;
                for (int b = 0; b < 16; ++b) y |= (uint64_t)(1u << ((x >> (2 * b)) & 3)) << (4 * b);
                uint32_t take = std::min<uint32_t>(8, (L + 1) / 2 - i);
                memcpy(p + i, &y, take);
            }
            p += (L + 1) / 2;
            for (uint32_t i = 0; i < L; i += 8) {
                uint64_t x = r.next(); uint8_t q[8];
                for (int b = 0; b < 8; ++b) { uint32_t v = (x >> (8 * b)) & 0xff; q[b] = qual8 ? (uint8_t)(2 + 5 * (v & 7)) : (uint8_t)(2 + (v * 40 >> 8)); }
                memcpy(p + i, q, std::min<uint32_t>(8, L - i));
            }
            p += L;
            *p++ = 'N'; *p++ = 'H'; *p++ = 'C'; *p++ = 1;
            if (rc.xs) { *p++ = 'X'; *p++ = 'S'; *p++ = 'A'; *p++ = rc.xs; }
            if (n_barcodes) {                                  // --barcodes N: CB:Z on 97 % of the reads, skewed towards low indices
                const double u = r.uni();
                if (r.uni() < 0.97) {
                    uint64_t h = (uint64_t)((double)n_barcodes * u * u) * 0x9E3779B97F4A7C15ull + 0xD1B54A32D192ED03ull;
                    h ^= h >> 31; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 29;
                    *p++ = 'C'; *p++ = 'B'; *p++ = 'Z';
                    for (int b = 0; b < 16; ++b) *p++ = "ACGT"[(h >> (2 * b)) & 3];
                    *p++ = '-'; *p++ = '1'; *p++ = 0;
                }
            }
            int32_t block_size = (int32_t)(p - rec_buf - 4);
            memcpy(rec_buf, &block_size, 4);
            raw.insert(raw.end(), rec_buf, p);
            while (raw.size() >= BLOCK_IN) flush_block(BLOCK_IN);
            // virtual offset after this record: htslib's bgzf_tell reports (next block, 0) when the
            // block is exactly consumed; within a task blocks are cut at BLOCK_IN bytes
            t.idx.push_back(Task::IdxRec{rc.pos, end, coff_rel << 16 | (uint64_t)raw.size()});
        }
        if (!raw.empty()) {
            // close the task's last block: offsets recorded as (coff_rel, raw.size()) stay valid except the
            // very last one, which equals "end of block" -> fix up to (next block, 0) like bgzf_tell does
            uint32_t n = (uint32_t)raw.size();
            uint64_t last_coff = coff_rel;
            flush_block(n);
            for (size_t i = t.idx.size(); i-- > 0;) {
                if (t.idx[i].voff_after == (last_coff << 16 | n)) t.idx[i].voff_after = coff_rel << 16; else break;
            }
        }
        deflateEnd(&zs);
    }
};

void usage() {
    fprintf(stderr, "usage: bamgen gen --out x.bam --config c2|c3|tiny --reads N [--seed S] [--level 1-9] [--threads T] [--qual8] [--barcodes N]\n"
                    "       bamgen index x.bam\n");
    exit(2);
}

// ------------------------------------------------------------------------------------------ gen
int cmd_gen(int argc, char** argv) {
    std::string out, cfgname = "c2"; uint64_t reads = 1000000, seed = 1234; int level = 6, threads = 0; bool qual8 = false; uint64_t n_barcodes = 0;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) usage(); return argv[++i]; };
        if (a == "--out") out = val(); else if (a == "--config") cfgname = val();
        else if (a == "--reads") reads = strtoull(val(), nullptr, 10); else if (a == "--seed") seed = strtoull(val(), nullptr, 10);
        else if (a == "--level") level = atoi(val()); else if (a == "--threads") threads = atoi(val());
        else if (a == "--qual8") qual8 = true;
        else if (a == "--barcodes") n_barcodes = strtoull(val(), nullptr, 10);      // single-cell shape: CB:Z tags for `-b`
        else usage();
    }
    if (out.empty()) usage();
    if (threads <= 0) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    Gen g; g.cfg = make_config(cfgname); g.seed = seed; g.level = level; g.qual8 = qual8; g.n_barcodes = n_barcodes;
    const Config& cfg = g.cfg;
    const uint32_t L = cfg.read_len;
    // ---- catalog
    Rng cr(seed ^ 0xC0FFEEull);
    int64_t genome = 0; for (auto& c : cfg.contigs) genome += c.len;
    for (size_t t = 0; t < cfg.contigs.size(); ++t) {
        int64_t len = cfg.contigs[t].len;
        int64_t n = std::max<int64_t>(4, len / 12000);
        int64_t hi = len - 700000; if (hi < 2000) hi = std::max<int64_t>(len / 2, 1500);
        for (int64_t k = 0; k < n; ++k) {
            Junc j; j.tid = (int32_t)t;
            j.donor = 1000 + (int64_t)(cr.uni() * (double)(hi - 1000));
            double u = cr.uni();
            if (u < 0.0025) j.len = 20 + cr.below(50);
            else if (u < 0.005) j.len = 500001 + cr.below(100000);
            else j.len = (uint32_t)std::exp(std::log(70.0) + cr.uni() * (std::log(100000.0) - std::log(70.0)));
            if ((int64_t)j.donor + j.len + 25000 + 2 * L >= len) j.len = 70 + cr.below(200);
            j.strand = cr.uni() < 0.5 ? '+' : '-';
            j.exon2 = 20 + cr.below(std::min<uint32_t>(41, L / 3));
            j.len2 = (uint32_t)std::exp(std::log(70.0) + cr.uni() * (std::log(20000.0) - std::log(70.0)));
            j.n_reads = 0; j.id = 0;
            g.cat.push_back(j);
        }
    }
    // Zipf(1) weights over a random global ranking
    {
        size_t J = g.cat.size();
        std::vector<uint32_t> rank(J);
        for (size_t i = 0; i < J; ++i) rank[i] = (uint32_t)i;
        for (size_t i = J; i-- > 1;) std::swap(rank[i], rank[cr.below((uint32_t)i + 1)]);
        double W = 0; for (size_t i = 0; i < J; ++i) W += 1.0 / (double)(i + 1);
        double S = cfg.spliced * (double)reads;
        for (size_t i = 0; i < J; ++i) {
            double x = S / W / (double)(rank[i] + 1);
            g.cat[i].n_reads = (uint64_t)(x + cr.uni());
        }
    }
    std::stable_sort(g.cat.begin(), g.cat.end(), [](const Junc& a, const Junc& b) { return a.tid != b.tid ? a.tid < b.tid : a.donor < b.donor; });
    uint64_t spliced_total = 0;
    for (size_t i = 0; i < g.cat.size(); ++i) { g.cat[i].id = i; spliced_total += g.cat[i].n_reads; }
    uint64_t unspliced_total = reads > spliced_total ? reads - spliced_total : 0;
    // ---- tasks
    std::vector<Task> tasks;
    {
        size_t jlo = 0; uint64_t acc_len = 0;
        for (size_t t = 0; t < cfg.contigs.size(); ++t) {
            int64_t len = cfg.contigs[t].len;
            int64_t nw = (len + cfg.window - 1) / cfg.window;
            // unspliced reads of the contig proportional to its length (prefix rounding keeps the total exact)
            uint64_t u_lo = (uint64_t)((long double)unspliced_total * acc_len / genome);
            uint64_t u_hi = (uint64_t)((long double)unspliced_total * (acc_len + (uint64_t)len) / genome);
            acc_len += (uint64_t)len;
            uint64_t u_c = u_hi - u_lo;
            size_t j_end = jlo; while (j_end < g.cat.size() && g.cat[j_end].tid == (int32_t)t) ++j_end;
            for (int64_t w = 0; w < nw; ++w) {
                Task k; k.tid = (int32_t)t; k.ws = w * cfg.window; k.we = std::min(len, (w + 1) * cfg.window);
                k.n_unspliced = (uint64_t)((long double)u_c * (w + 1) / nw) - (uint64_t)((long double)u_c * w / nw);
                // junctions whose reads may start in this window: donor in [ws, we + L)
                size_t a = jlo; while (a < j_end && g.cat[a].donor < k.ws) ++a;
                size_t b = a; while (b < j_end && g.cat[b].donor < k.we + (int64_t)L + 16) ++b;
                k.j_lo = a; k.j_hi = b; k.id = tasks.size();
                tasks.push_back(std::move(k));
            }
            jlo = j_end;
        }
    }
    // ---- header
    std::vector<uint8_t> hdr;
    {
        std::string text = "@HD\tVN:1.4\tSO:coordinate\n";
        for (auto& c : cfg.contigs) text += "@SQ\tSN:" + c.name + "\tLN:" + std::to_string(c.len) + "\n";
        text += "@PG\tID:bamgen\tPN:bamgen\n";
        std::vector<uint8_t> raw;
        auto put32 = [&](int32_t v) { uint8_t b[4]; memcpy(b, &v, 4); raw.insert(raw.end(), b, b + 4); };
        raw.insert(raw.end(), {'B', 'A', 'M', 1});
        put32((int32_t)text.size()); raw.insert(raw.end(), text.begin(), text.end());
        put32((int32_t)cfg.contigs.size());
        for (auto& c : cfg.contigs) { put32((int32_t)c.name.size() + 1); raw.insert(raw.end(), c.name.begin(), c.name.end()); raw.push_back(0); put32((int32_t)c.len); }
        z_stream zs; memset(&zs, 0, sizeof zs);
        deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
        for (size_t o = 0; o < raw.size(); o += BLOCK_IN) bgzf_block(zs, raw.data() + o, (uint32_t)std::min<size_t>(BLOCK_IN, raw.size() - o), level, hdr);
        deflateEnd(&zs);
    }
    FILE* f = fopen(out.c_str(), "wb");
    if (!f) { perror("open output"); return 1; }
    fwrite(hdr.data(), 1, hdr.size(), f);
    uint64_t base = hdr.size();
    BaiBuilder bai((int)cfg.contigs.size());
    bai.set_start(base << 16);
    // ---- workers + ordered writer
    std::mutex mu; std::condition_variable cv;
    size_t next_task = 0, next_write = 0;
    const size_t ahead = (size_t)threads * 3 + 2;
    std::vector<std::thread> pool;
    for (int w = 0; w < threads; ++w) pool.emplace_back([&] {
        for (;;) {
            size_t i;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return next_task >= tasks.size() || next_task < next_write + ahead; });
                if (next_task >= tasks.size()) return;
                i = next_task++;
            }
            g.run_task(tasks[i]);
            { std::lock_guard<std::mutex> lk(mu); tasks[i].done = true; }
            cv.notify_all();
        }
    });
    uint64_t written = 0;
    for (size_t i = 0; i < tasks.size(); ++i) {
        { std::unique_lock<std::mutex> lk(mu); cv.wait(lk, [&] { return tasks[i].done; }); }
        Task& t = tasks[i];
        fwrite(t.comp.data(), 1, t.comp.size(), f);
        for (const auto& r : t.idx) bai.push(t.tid, r.beg, r.end, (((r.voff_after >> 16) + base) << 16) | (r.voff_after & 0xffff), true);
        written += t.idx.size();
        base += t.comp.size();
        std::vector<uint8_t>().swap(t.comp); std::vector<Task::IdxRec>().swap(t.idx);
        { std::lock_guard<std::mutex> lk(mu); next_write = i + 1; }
        cv.notify_all();
    }
    for (auto& th : pool) th.join();
    bai.finish(base << 16);
    fwrite(EOF_BLOCK, 1, sizeof EOF_BLOCK, f);
    fclose(f);
    if (!bai.save(out + ".bai")) { perror("write bai"); return 1; }
    fprintf(stderr, "bamgen: %llu reads (%llu spliced) %zu catalog junctions -> %s (%.1f MB)\n", (unsigned long long)written,
            (unsigned long long)spliced_total, g.cat.size(), out.c_str(), (double)(base + 28) / 1e6);
    printf("{\"reads\": %llu, \"bytes\": %llu}\n", (unsigned long long)written, (unsigned long long)(base + 28));
    return 0;
}

// ------------------------------------------------------------------------------------------ index
int cmd_index(int argc, char** argv) {
    if (argc < 3) usage();
    std::string path = argv[2];
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { perror("open"); return 1; }
    std::vector<uint8_t> cbuf(0x10000), ubuf;      // ubuf: sliding window of inflated bytes
    struct Blk { uint64_t coff; uint32_t ulen; };
    std::vector<Blk> blks;                           // blocks currently (partly) in ubuf
    size_t upos = 0;                                 // consumed bytes in ubuf (relative to blks[0] start)
    uint64_t coff = 0; bool eof = false;
    z_stream zs; memset(&zs, 0, sizeof zs); inflateInit2(&zs, -15);
    auto more = [&]() -> bool {
        if (eof) return false;
        uint8_t h[18];
        if (fread(h, 1, 18, f) != 18) { eof = true; return false; }
        uint32_t bsize = (uint32_t)(h[16] | h[17] << 8) + 1;
        if (fread(cbuf.data(), 1, bsize - 18, f) != bsize - 18) { eof = true; return false; }
        uint32_t isize; memcpy(&isize, cbuf.data() + bsize - 18 - 4, 4);
        size_t o = ubuf.size(); ubuf.resize(o + isize);
        inflateReset(&zs);
        zs.next_in = cbuf.data(); zs.avail_in = bsize - 18 - 8; zs.next_out = ubuf.data() + o; zs.avail_out = isize;
        if (isize && inflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "inflate failed at %llu\n", (unsigned long long)coff); exit(1); }
        blks.push_back(Blk{coff, isize});
        coff += bsize;
        return true;
    };
    auto need = [&](size_t n) -> bool { while (ubuf.size() - upos < n) if (!more()) return false; return true; };
    auto tell = [&]() -> uint64_t {                  // virtual offset of upos, bgzf_tell style
        size_t o = upos;
        for (size_t i = 0; i < blks.size(); ++i) {
            if (o < blks[i].ulen || (o == 0 && blks[i].ulen == 0)) return blks[i].coff << 16 | o;
            o -= blks[i].ulen;
        }
        return coff << 16;                           // exactly at the end of everything read so far
    };
    auto trim = [&]() {                              // drop fully consumed blocks
        size_t drop = 0, nb = 0;
        while (nb < blks.size() && drop + blks[nb].ulen <= upos) { drop += blks[nb].ulen; ++nb; }
        if (nb) { ubuf.erase(ubuf.begin(), ubuf.begin() + drop); upos -= drop; blks.erase(blks.begin(), blks.begin() + nb); }
    };
    // header
    if (!need(12) || memcmp(ubuf.data(), "BAM\1", 4)) { fprintf(stderr, "not a BAM\n"); return 1; }
    int32_t l_text; memcpy(&l_text, ubuf.data() + 4, 4);
    if (!need(12 + (size_t)l_text)) return 1;
    int32_t n_ref; memcpy(&n_ref, ubuf.data() + 8 + l_text, 4);
    upos = 12 + (size_t)l_text;
    for (int32_t i = 0; i < n_ref; ++i) {
        if (!need(4)) return 1;
        int32_t l_name; memcpy(&l_name, ubuf.data() + upos, 4);
        if (!need(8 + (size_t)l_name)) return 1;
        upos += 8 + (size_t)l_name;
    }
    trim();
    BaiBuilder bai(n_ref);
    // make sure the block holding the first record is loaded so tell() is exact
    if (!need(1)) { bai.finish(coff << 16); bai.save(path + ".bai"); return 0; }
    bai.set_start(tell());
    uint64_t n = 0;
    while (need(4)) {
        int32_t bs; memcpy(&bs, ubuf.data() + upos, 4);
        if (bs < 32 || !need(4 + (size_t)bs)) break;
        const uint8_t* c = ubuf.data() + upos + 4;
        int32_t tid, pos; memcpy(&tid, c, 4); memcpy(&pos, c + 4, 4);
        uint32_t l_qname = c[8]; uint16_t n_cigar, flag; memcpy(&n_cigar, c + 12, 2); memcpy(&flag, c + 14, 2);
        int32_t end = pos + 1;
        if (!(flag & 4) && n_cigar) {
            end = pos;
            for (uint32_t k = 0; k < n_cigar; ++k) { uint32_t w; memcpy(&w, c + 32 + l_qname + 4 * k, 4); uint32_t op = w & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) end += (int32_t)(w >> 4); }
            if (end == pos) end = pos + 1;            // bam_endpos (sam.c:336-342): a CIGAR that consumes no reference counts as 1
        }
        upos += 4 + (size_t)bs;
        // bgzf_tell after the record: if the block is exactly consumed htslib reports (next block, 0),
        // which needs the next block's address: tell() returns `coff` of the next unread block then.
        uint64_t after = tell();
        bai.push(tid, pos, end, after, !(flag & 4));
        ++n;
        trim();
    }
    inflateEnd(&zs); fclose(f);
    bai.finish(tell());
    if (!bai.save(path + ".bai")) { perror("write bai"); return 1; }
    fprintf(stderr, "bamgen index: %llu records\n", (unsigned long long)n);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) usage();
    if (!strcmp(argv[1], "gen")) return cmd_gen(argc, argv);
    if (!strcmp(argv[1], "index")) return cmd_index(argc, argv);
    usage();
    return 2;
}
