// regtools_b200/csrc/bam_feeder.cc — see bam_feeder.h.
#include "bam_feeder.h"

#include <zlib.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <unordered_map>

namespace rtjx {

namespace {

inline uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline int32_t rdi32(const uint8_t* p) { int32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- BGZF block header (bgzf.c:348-355 check_header, :525-546) ---------------------------------
struct BlockDesc {
    uint64_t coff;      // compressed offset of the block
    uint32_t csize;     // BSIZE (whole block)
    uint32_t isize;     // ISIZE from the trailer
};

// 0 ok, 1 clean end of file, -1 malformed / truncated
int peek_block(const uint8_t* file, size_t size, uint64_t coff, BlockDesc* d) {
    if (coff >= size) return 1;
    if (coff + 18 > size) return -1;
    const uint8_t* h = file + coff;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return -1;
    if (rd16(h + 10) != 6 || h[12] != 'B' || h[13] != 'C' || rd16(h + 14) != 2) return -1;
    uint32_t bsize = (uint32_t)rd16(h + 16) + 1u;
    if (bsize < 26 || coff + bsize > size) return -1;
    d->coff = coff; d->csize = bsize; d->isize = rd32(h + bsize - 4);
    return 0;
}

// One reusable raw-deflate decoder per thread (the reference does inflateInit2/inflateEnd per block,
// bgzf.c:292-316).
class Inflater {
public:
    Inflater() { memset(&zs_, 0, sizeof zs_); ok_ = inflateInit2(&zs_, -15) == Z_OK; }
    ~Inflater() { if (ok_) inflateEnd(&zs_); }
    // returns inflated length or -1
    long run(const uint8_t* file, const BlockDesc& b, uint8_t* dst, uint32_t cap) {
        if (!ok_ || inflateReset(&zs_) != Z_OK) return -1;
        // inflate_block (bgzf.c:292-316) hands zlib `block_length - 16` bytes: the deflate data, the 8-byte trailer and two
        // bytes of whatever its buffer held beyond the block.  A well-formed stream ends before the trailer either way; a
        // damaged one may run on into it, so the trailer is offered here too (the two stray bytes are not reproducible).
        zs_.next_in = const_cast<Bytef*>(file + b.coff + 18);
        zs_.avail_in = b.csize - 18;
        zs_.next_out = dst; zs_.avail_out = cap;
        int rc = inflate(&zs_, Z_FINISH);
        if (rc != Z_STREAM_END) return -1;
        return (long)zs_.total_out;
    }
private:
    z_stream zs_; bool ok_;
};

// ---- sequential reader with htslib's exact bgzf_read / seek / tell behaviour ---------------------
// Used for the header, the BAI-chunked region path and nothing hot.
class SeqBgzf {
public:
    SeqBgzf(const uint8_t* file, size_t size) : file_(file), size_(size) {}
    int read_block() {                                   // bgzf.c:421-546
        BlockDesc d;
        int rc = peek_block(file_, size_, fpos_, &d);
        if (rc == 1) { block_length_ = 0; return 0; }
        if (rc < 0) return -1;
        long n = inf_.run(file_, d, buf_, sizeof buf_);
        if (n < 0) return -1;
        fpos_ = d.coff + d.csize;
        if (block_length_ != 0) block_offset_ = 0;
        block_address_ = d.coff;
        block_length_ = (int32_t)n;
        return 0;
    }
    long read(void* dst, size_t n) {                     // bgzf.c:548-577
        size_t got = 0;
        while (got < n) {
            int32_t avail = block_length_ - block_offset_;
            if (avail <= 0) {
                if (read_block() != 0) return -1;
                avail = block_length_ - block_offset_;
                if (avail <= 0) break;                   // empty block == end of data for this call
            }
            size_t take = std::min(n - got, (size_t)avail);
            memcpy((uint8_t*)dst + got, buf_ + block_offset_, take);
            block_offset_ += (int32_t)take; got += take;
        }
        if (block_offset_ == block_length_) { block_address_ = fpos_; block_offset_ = block_length_ = 0; }
        return (long)got;
    }
    void seek(uint64_t voff) {                           // bgzf.c:848-867
        fpos_ = voff >> 16; block_length_ = 0; block_address_ = voff >> 16; block_offset_ = (int32_t)(voff & 0xffff);
    }
    uint64_t tell() const { return block_address_ << 16 | ((uint64_t)block_offset_ & 0xffff); }
private:
    const uint8_t* file_; size_t size_;
    uint64_t fpos_ = 0, block_address_ = 0;
    int32_t block_length_ = 0, block_offset_ = 0;
    Inflater inf_;
    uint8_t buf_[0x10000];
};

// ---- aux scan: bam_aux_get + bam_aux2A (sam.c:1254-1266, 1301-1307, 1233-1252) -------------------
inline uint8_t strand_tag_byte(const uint8_t* s, const uint8_t* e, const char tag[2]) {
    while (s + 3 <= e) {
        const bool match = s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1];
        const uint8_t type = s[2];
        s += 3;
        if (match) return (type == 'A' && s < e) ? *s : 0;
        switch (type) {
        case 'A': case 'c': case 'C': s += 1; break;
        case 's': case 'S': s += 2; break;
        case 'i': case 'I': case 'f': s += 4; break;
        case 'd': s += 8; break;
        case 'Z': case 'H': {
            const void* z = memchr(s, 0, (size_t)(e - s));
            if (!z) return 0;
            s = (const uint8_t*)z + 1; break; }
        case 'B': {
            if (s + 5 > e) return 0;
            uint8_t sub = s[0]; uint32_t n = rd32(s + 1); s += 5;
            size_t sz = (sub == 'c' || sub == 'C' || sub == 'A') ? 1 : (sub == 's' || sub == 'S') ? 2 :
                        (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : sub == 'd' ? 8 : 0;
            if ((size_t)(e - s) < sz * n) return 0;
            s += sz * n; break; }
        default: return 0;     // the reference abort()s on an unknown type (sam.c:1246-1247)
        }
    }
    return 0;
}

// ---- barcode tag: bam_aux_get + bam_aux2Z (sam.c:1254-1266, 1309-1315) ------------------------------
// 0 = tag absent, 1 = Z/H value at [*val, *val + *len), 2 = tag present with another type
inline int barcode_tag_value(const uint8_t* s, const uint8_t* e, const char tag[2], const uint8_t** val, size_t* len) {
    while (s + 3 <= e) {
        const bool match = s[0] == (uint8_t)tag[0] && s[1] == (uint8_t)tag[1];
        const uint8_t type = s[2];
        s += 3;
        if (match) {
            if (type != 'Z' && type != 'H') return 2;
            const void* z = memchr(s, 0, (size_t)(e - s));
            *val = s; *len = z ? (size_t)((const uint8_t*)z - s) : (size_t)(e - s);
            return 1;
        }
        switch (type) {
        case 'A': case 'c': case 'C': s += 1; break;
        case 's': case 'S': s += 2; break;
        case 'i': case 'I': case 'f': s += 4; break;
        case 'd': s += 8; break;
        case 'Z': case 'H': {
            const void* z = memchr(s, 0, (size_t)(e - s));
            if (!z) return 0;
            s = (const uint8_t*)z + 1; break; }
        case 'B': {
            if (s + 5 > e) return 0;
            uint8_t sub = s[0]; uint32_t n = rd32(s + 1); s += 5;
            size_t sz = (sub == 'c' || sub == 'C' || sub == 'A') ? 1 : (sub == 's' || sub == 'S') ? 2 :
                        (sub == 'i' || sub == 'I' || sub == 'f') ? 4 : sub == 'd' ? 8 : 0;
            if ((size_t)(e - s) < sz * n) return 0;
            s += sz * n; break; }
        default: return 0;     // the reference abort()s on an unknown type (sam.c:1246-1247)
        }
    }
    return 0;
}

// ---- batch writer --------------------------------------------------------------------------------
class BatchWriter {
public:
    BatchWriter(BatchSink* sink, const FeederOptions& opt, int32_t n_ref)
        : sink_(sink), opt_(opt), n_ref_(n_ref) {}
    // rec points at the 32-byte core (after block_size); l_data = block_size - 32 already validated
    inline void push(const uint8_t* core, int32_t l_data) {
        const uint32_t n_cigar = rd16(core + 12);
        if (!cur_ || cur_->n_reads == cur_->cap_reads || cur_->n_ops + n_cigar > cur_->cap_ops) roll();
        HostBatch& b = *cur_;
        const uint32_t i = b.n_reads;
        int32_t tid = rdi32(core);
        const uint32_t l_qname = core[8], mapq = core[9], flag = rd16(core + 14);
        const uint8_t* data = core + 32;
        const uint8_t* cig = data + l_qname;
        uint32_t strand = 0, bc = 0;
        if (n_cigar > 1) {
            if (tid < 0 || tid >= n_ref_) tid = -1;
            uint32_t* dst = b.cigar + b.n_ops;
            uint32_t nn = 0;
            for (uint32_t k = 0; k < n_cigar; ++k) { uint32_t w = rd32(cig + 4 * k); dst[k] = w; nn += (w & 0xf) == 3; }
            b.n_junction_ops += nn;
            if (opt_.xs_mode && nn) {
                const int32_t l_qseq = rdi32(core + 16);
                const uint8_t* aux = cig + 4 * (size_t)n_cigar + ((size_t)l_qseq + 1) / 2 + (size_t)l_qseq;
                strand = strand_tag_byte(aux, data + l_data, opt_.tag);
            }
            if (opt_.barcodes) {                               // set_junction_barcode is called for every n_cigar > 1 alignment
                const int32_t l_qseq = rdi32(core + 16);
                const uint8_t* aux = cig + 4 * (size_t)n_cigar + ((size_t)l_qseq + 1) / 2 + (size_t)l_qseq;
                const uint8_t* val = nullptr; size_t len = 0;
                const int k = barcode_tag_value(aux, data + l_data, opt_.bc_tag, &val, &len);
                if (k == 1) bc = opt_.barcodes->intern((const char*)val, len);
                else {
                    if (k == 0) ++opt_.barcodes->missing; else ++opt_.barcodes->bad_type;
                    bc = opt_.barcodes->intern("?", 1);
                }
            }
        } else if (n_cigar == 1) {
            b.cigar[b.n_ops] = rd32(cig);
        }
        b.tid[i] = tid;
        b.pos[i] = rdi32(core + 4);
        b.meta[i] = flag << 16 | mapq << 8 | strand;
        b.cig_off[i] = b.n_ops;
        if (opt_.barcodes) b.bc[i] = bc;
        b.n_ops += n_cigar;
        b.n_reads = i + 1;
        ++total_reads_;
        total_ops_ += n_cigar;
    }
    void finish() {
        if (cur_) { cur_->cig_off[cur_->n_reads] = cur_->n_ops; sink_->submit(cur_); cur_ = nullptr; }
    }
    uint64_t total_reads() const { return total_reads_; }
    uint64_t total_ops() const { return total_ops_; }
private:
    void roll() {
        if (cur_) { cur_->cig_off[cur_->n_reads] = cur_->n_ops; sink_->submit(cur_); }
        cur_ = sink_->acquire();
        cur_->n_reads = cur_->n_ops = cur_->n_junction_ops = 0;
        cur_->first_ordinal = total_reads_;
    }
    BatchSink* sink_; FeederOptions opt_; int32_t n_ref_;
    HostBatch* cur_ = nullptr;
    uint64_t total_reads_ = 0, total_ops_ = 0;
};

}  // namespace
struct BarcodeDict::Map { std::unordered_map<std::string, uint32_t> ids; std::string last; uint32_t last_id = 0; bool have_last = false; };
BarcodeDict::BarcodeDict() : map_(new Map()) {}
BarcodeDict::~BarcodeDict() { delete map_; }
void BarcodeDict::clear() { names.clear(); missing = bad_type = 0; map_->ids.clear(); map_->have_last = false; }
uint32_t BarcodeDict::intern(const char* s, size_t n) {
    Map& m = *map_;
    if (m.have_last && m.last.size() == n && memcmp(m.last.data(), s, n) == 0) return m.last_id;
    m.last.assign(s, n);
    auto it = m.ids.find(m.last);
    if (it == m.ids.end()) { it = m.ids.emplace(m.last, (uint32_t)names.size()).first; names.push_back(m.last); }
    m.last_id = it->second; m.have_last = true;
    return m.last_id;
}
namespace {

// bam_read1's validity checks (sam.c:399-432).  core = 32 bytes after block_size.
inline bool record_ok(const uint8_t* core, int32_t block_len) {
    const int32_t l_data = block_len - 32;
    const int32_t l_qseq = rdi32(core + 16);
    const uint32_t l_qname = core[8];
    if (l_data < 0 || l_qseq < 0 || l_qname < 1) return false;
    const int64_t aux_off = (int64_t)l_qname + 4ll * rd16(core + 12) + ((int64_t)l_qseq + 1) / 2 + l_qseq;
    return aux_off <= l_data;
}

// bam_endpos (sam.c:336-342); cigar type bit 2 = consumes reference (M D N = X)
inline int32_t record_endpos(const uint8_t* core) {
    const uint32_t n_cigar = rd16(core + 12), flag = rd16(core + 14);
    const int32_t pos = rdi32(core + 4);
    if (!(flag & 4u) && n_cigar > 0) {
        const uint8_t* cig = core + 32 + core[8];
        int32_t l = 0;
        for (uint32_t k = 0; k < n_cigar; ++k) {
            uint32_t w = rd32(cig + 4 * k);
            if ((0x3C1A7u >> ((w & 0xf) << 1) & 3u) & 2u) l += (int32_t)(w >> 4);
        }
        return pos + l;
    }
    return pos + 1;
}

// ---- region path: hts_itr_next over BAI chunks (hts.c:1941-1963), single threaded ----------------
void feed_region(const BamFile& bam, const std::vector<Chunk64>& off, const IterSpec& spec, BatchWriter& w,
                 FeederStats* st) {
    if (off.empty()) return;
    const double t_begin = now_s();
    SeqBgzf fp(bam.data(), bam.size());
    std::vector<uint8_t> rec;
    long long ci = -1; uint64_t curr_off = 0;
    for (;;) {
        if (curr_off == 0 || curr_off >= off[ci].end) {
            if (ci == (long long)off.size() - 1) break;
            if (ci < 0 || off[ci].end != off[ci + 1].beg) { fp.seek(off[ci + 1].beg); curr_off = fp.tell(); }
            ++ci;
        }
        int32_t block_len;
        if (fp.read(&block_len, 4) != 4) break;
        if (block_len < 32) break;
        rec.resize((size_t)block_len);
        if (fp.read(rec.data(), 32) != 32) break;
        if (!record_ok(rec.data(), block_len)) break;
        if (fp.read(rec.data() + 32, (size_t)block_len - 32) != (long)block_len - 32) break;
        curr_off = fp.tell();
        const int32_t tid = rdi32(rec.data()), pos = rdi32(rec.data() + 4);
        if (tid != spec.tid || pos >= spec.end) break;
        if (record_endpos(rec.data()) > spec.beg && spec.end > pos) w.push(rec.data(), block_len - 32);
        if (st) st->inflated_bytes += 4 + (uint64_t)block_len;
    }
    if (st) st->parse_s += now_s() - t_begin;            // (inflate and record split alternate on this one thread)
}

// ---- parallel stream: [voffset, EOF or end voffset) ranges -----------------------------------------
constexpr uint32_t CHUNK_BLOCKS = 32;

struct StreamChunk {
    std::vector<BlockDesc> blocks;
    std::vector<uint32_t> out_off;       // placement of each block in buf: one 64 KB slot per block (the trailer's ISIZE is
                                         // not trusted: htslib takes the block's length from zlib's total_out, bgzf.c:315)
    std::vector<int32_t> out_len;        // actual inflated length (-1: failed)
    std::vector<uint8_t> buf;
    bool ready = false;
    bool end_of_range = false;           // no further chunk belongs to the current range
    bool stream_ends = false;            // the block after this chunk is missing / malformed / past EOF: bgzf_read_block fails or
                                         // reads nothing there and hts_itr_next finishes the WHOLE iteration (hts.c:1928-1963)
    int range = 0;
    double inflate_s = 0;
};

class ParallelStream {
public:
    ParallelStream(const BamFile& bam, std::vector<Chunk64> ranges, int n_threads)
        : bam_(bam), ranges_(std::move(ranges)) {
        n_threads_ = n_threads > 0 ? n_threads : (int)std::max(1u, std::thread::hardware_concurrency());
        ring_.resize((size_t)n_threads_ * 3 + 2);
        for (auto& c : ring_) c.reset(new StreamChunk());
        if (!ranges_.empty()) next_coff_ = ranges_[0].beg >> 16;
        for (int i = 0; i < n_threads_; ++i) workers_.emplace_back([this] { worker(); });
    }
    ~ParallelStream() {
        { std::lock_guard<std::mutex> g(mu_); stop_ = true; }
        cv_work_.notify_all();
        for (auto& t : workers_) t.join();
    }
    // Next chunk in order, or nullptr when every range is exhausted.  Caller must release().
    StreamChunk* next(double* wait_s) {
        top_up();
        if (head_ == tail_) return nullptr;
        StreamChunk* c = ring_[head_ % ring_.size()].get();
        double t0 = now_s();
        { std::unique_lock<std::mutex> g(mu_); cv_done_.wait(g, [&] { return c->ready; }); }
        if (wait_s) *wait_s += now_s() - t0;
        return c;
    }
    void release() { ++head_; }
    // abandon the rest of the current range (end voffset reached) and move to the next one
    void skip_to_range(int r) {
        if (sched_range_ < r) {
            sched_range_ = r;
            sched_done_ = sched_range_ >= (int)ranges_.size();
            if (!sched_done_) next_coff_ = ranges_[sched_range_].beg >> 16;
        }
    }
    void cancel() { sched_done_ = true; }
    uint64_t blocks = 0, cbytes = 0;
    double inflate_s() {
        double s = 0; for (auto& c : ring_) s += c->inflate_s; return s;
    }
private:
    // Walk block headers of the current range and queue chunks while ring slots are free.
    void top_up() {
        while (!sched_done_ && tail_ - head_ < ring_.size()) {
            StreamChunk* c = ring_[tail_ % ring_.size()].get();
            c->blocks.clear(); c->out_off.clear(); c->out_len.clear();
            c->ready = false; c->end_of_range = false; c->stream_ends = false; c->range = sched_range_;
            const uint64_t end_coff = ranges_[sched_range_].end == UINT64_MAX ? UINT64_MAX : ranges_[sched_range_].end >> 16;
            const bool end_has_tail = ranges_[sched_range_].end != UINT64_MAX && (ranges_[sched_range_].end & 0xffff) != 0;
            uint32_t total = 0;
            while (c->blocks.size() < CHUNK_BLOCKS) {
                if (next_coff_ > end_coff || (next_coff_ == end_coff && !end_has_tail)) { c->end_of_range = true; break; }
                BlockDesc d;
                int rc = peek_block(bam_.data(), bam_.size(), next_coff_, &d);
                if (rc != 0) { c->end_of_range = true; c->stream_ends = true; break; }   // end of file or a bad header: reading ends
                c->blocks.push_back(d); c->out_off.push_back(total); c->out_len.push_back(-1);
                total += 0x10000u;
                next_coff_ = d.coff + d.csize;
                blocks++; cbytes += d.csize;
            }
            if (c->buf.size() < total) c->buf.resize(total);
            if (c->end_of_range) {
                ++sched_range_;
                sched_done_ = sched_range_ >= (int)ranges_.size();
                if (!sched_done_) next_coff_ = ranges_[sched_range_].beg >> 16;
            }
            if (c->blocks.empty() && !c->end_of_range) continue;
            {
                std::lock_guard<std::mutex> g(mu_);
                queue_.push_back(c);
            }
            ++tail_;
            cv_work_.notify_one();
        }
    }
    void worker() {
        Inflater inf;
        for (;;) {
            StreamChunk* c;
            {
                std::unique_lock<std::mutex> g(mu_);
                cv_work_.wait(g, [&] { return stop_ || !queue_.empty(); });
                if (stop_) return;
                c = queue_.front(); queue_.pop_front();
            }
            double t0 = now_s();
            for (size_t i = 0; i < c->blocks.size(); ++i)
                c->out_len[i] = (int32_t)inf.run(bam_.data(), c->blocks[i], c->buf.data() + c->out_off[i], 0x10000u);
            c->inflate_s += now_s() - t0;
            { std::lock_guard<std::mutex> g(mu_); c->ready = true; }
            cv_done_.notify_all();
        }
    }
    const BamFile& bam_;
    std::vector<Chunk64> ranges_;
    int n_threads_;
    std::vector<std::unique_ptr<StreamChunk>> ring_;
    size_t head_ = 0, tail_ = 0;
    uint64_t next_coff_ = 0; int sched_range_ = 0; bool sched_done_ = false;
    std::mutex mu_; std::condition_variable cv_work_, cv_done_;
    std::deque<StreamChunk*> queue_;
    std::vector<std::thread> workers_;
    bool stop_ = false;
};

// Record splitter over the ordered chunk stream of one or more ranges.
void feed_stream(const BamFile& bam, const std::vector<Chunk64>& ranges, const FeederOptions& opt, BatchWriter& w,
                 FeederStats* st) {
    if (ranges.empty()) return;
    ParallelStream ps(bam, ranges, opt.n_threads);
    std::vector<uint8_t> carry;          // partial record spilling over a block boundary
    int cur_range = 0, done_range = -1;
    uint32_t skip = (uint32_t)(ranges[0].beg & 0xffff);   // bytes to skip in the first block of the range
    bool fresh_range = true, dead = false;
    double parse_s = 0, wait_s = 0;
    StreamChunk* c;
    while (!dead && (c = ps.next(&wait_s)) != nullptr) {
        double t0 = now_s();
        if (c->range <= done_range) { ps.release(); continue; }   // leftovers of a range that already ended
        if (c->range != cur_range) {      // a new range starts: drop any partial record of the previous one
            cur_range = c->range; carry.clear(); fresh_range = true;
            skip = (uint32_t)(ranges[cur_range].beg & 0xffff);
        }
        const uint64_t end_voff = ranges[cur_range].end;
        bool range_done = false;
        for (size_t bi = 0; bi < c->blocks.size() && !dead && !range_done; ++bi) {
            // inflate failed (bgzf_read_block < 0) or the block is empty (bgzf_read breaks, bam_read1 reports the end of the
            // file, bgzf.c:559-561): the stream ends here, whatever the trailer's ISIZE says
            if (c->out_len[bi] <= 0) { dead = true; break; }
            const uint8_t* p = c->buf.data() + c->out_off[bi];
            const uint8_t* const bstart = p;
            const uint8_t* e = p + c->out_len[bi];
            const uint64_t bvoff = c->blocks[bi].coff << 16;
            if (st) st->inflated_bytes += (uint64_t)c->out_len[bi];
            if (fresh_range) { p += std::min<size_t>(skip, (size_t)(e - p)); fresh_range = false; }
            // finish a record carried over from the previous block(s)
            if (!carry.empty()) {
                while (carry.size() < 4 && p < e) carry.push_back(*p++);
                if (carry.size() < 4) continue;
                const int32_t bl = rdi32(carry.data());
                if (bl < 32) { dead = true; break; }
                const size_t need = 4 + (size_t)bl;
                const size_t take = std::min(need - carry.size(), (size_t)(e - p));
                carry.insert(carry.end(), p, p + take); p += take;
                if (carry.size() < need) continue;
                if (!record_ok(carry.data() + 4, bl)) { dead = true; break; }
                w.push(carry.data() + 4, bl - 32);
                carry.clear();
            }
            // whole records inside this block
            while (p < e) {
                if (end_voff != UINT64_MAX && (bvoff | (uint64_t)(p - bstart)) >= end_voff) { range_done = true; break; }
                if (e - p < 4) { carry.assign(p, e); p = e; break; }
                const int32_t bl = rdi32(p);
                if (bl < 32) { dead = true; break; }
                if ((size_t)(e - p) < 4 + (size_t)bl) { carry.assign(p, e); p = e; break; }
                if (!record_ok(p + 4, bl)) { dead = true; break; }
                w.push(p + 4, bl - 32);
                p += 4 + (size_t)bl;
            }
        }
        if (c->stream_ends && !range_done) dead = true;
        const bool eor = c->end_of_range;
        const int r = c->range;
        ps.release();
        if (range_done) { ps.skip_to_range(r + 1); carry.clear(); done_range = r; }
        if (eor) carry.clear();            // truncated record at the end of a range is dropped (bam_read1 < 0)
        parse_s += now_s() - t0;
        if (dead) ps.cancel();
    }
    if (st) {
        st->bgzf_blocks += ps.blocks; st->compressed_bytes += ps.cbytes;
        st->inflate_s += ps.inflate_s(); st->parse_s += parse_s; st->wait_s += wait_s;
    }
}

}  // namespace

// ---- BamFile -------------------------------------------------------------------------------------
BamFile::~BamFile() {
    if (map_ && size_) munmap(const_cast<uint8_t*>(map_), size_);
    if (fd_ >= 0) close(fd_);
}

bool BamFile::open(const std::string& path, std::string* err) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) { if (err) *err = "cannot open " + path; return false; }
    struct stat st;
    if (fstat(fd_, &st) != 0 || !S_ISREG(st.st_mode)) { if (err) *err = "not a regular file: " + path; return false; }
    size_ = (size_t)st.st_size;
    if (size_ == 0) { if (err) *err = "empty file: " + path; return false; }
    void* m = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd_, 0);
    if (m == MAP_FAILED) { map_ = nullptr; if (err) *err = "mmap failed: " + path; return false; }
    map_ = static_cast<const uint8_t*>(m);
    madvise(m, size_, MADV_SEQUENTIAL);
    return read_header(err);
}

// bam_hdr_read (sam.c:114-223)
bool BamFile::read_header(std::string* err) {
    SeqBgzf fp(map_, size_);
    char magic[4]; int32_t l_text = 0, n_ref = 0;
    if (fp.read(magic, 4) != 4 || memcmp(magic, "BAM\1", 4) != 0) { if (err) *err = "invalid BAM binary header"; return false; }
    if (fp.read(&l_text, 4) != 4 || l_text < 0) { if (err) *err = "truncated BAM header"; return false; }
    std::vector<char> text((size_t)l_text);
    if (l_text && fp.read(text.data(), (size_t)l_text) != l_text) { if (err) *err = "truncated BAM header"; return false; }
    if (fp.read(&n_ref, 4) != 4 || n_ref < 0) { if (err) *err = "truncated BAM header"; return false; }
    hdr_.names.resize((size_t)n_ref); hdr_.lengths.resize((size_t)n_ref);
    for (int32_t i = 0; i < n_ref; ++i) {
        int32_t l_name = 0;
        if (fp.read(&l_name, 4) != 4 || l_name < 0) { if (err) *err = "truncated BAM header"; return false; }
        std::vector<char> name((size_t)l_name + 1, 0);
        if (l_name && fp.read(name.data(), (size_t)l_name) != l_name) { if (err) *err = "truncated BAM header"; return false; }
        hdr_.names[i] = std::string(name.data());
        if (fp.read(&hdr_.lengths[i], 4) != 4) { if (err) *err = "truncated BAM header"; return false; }
    }
    hdr_.first_record_voffset = fp.tell();
    return true;
}

int32_t BamFile::name2id(const std::string& name) const {
    int32_t id = -1;
    for (size_t i = 0; i < hdr_.names.size(); ++i) if (hdr_.names[i] == name) id = (int32_t)i;
    return id;
}

// ---- BAI -----------------------------------------------------------------------------------------
const BaiIndex::Bin* BaiIndex::Ref::find(uint32_t bin) const {
    auto it = std::lower_bound(bins.begin(), bins.end(), bin, [](const Bin& b, uint32_t v) { return b.bin < v; });
    return (it != bins.end() && it->bin == bin) ? &*it : nullptr;
}

bool BaiIndex::load(const std::string& path) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0) { close(fd); return false; }
    std::vector<uint8_t> buf((size_t)st.st_size);
    size_t got = 0;
    while (got < buf.size()) {
        ssize_t r = ::read(fd, buf.data() + got, buf.size() - got);
        if (r <= 0) break;
        got += (size_t)r;
    }
    close(fd);
    if (got != buf.size() || buf.size() < 8) return false;
    // hts_idx_load_local reads the index through bgzf_open (hts.c:1575): a .csi is BGZF-compressed, a .bai is plain
    if (buf[0] == 0x1f && buf[1] == 0x8b) {
        std::vector<uint8_t> raw;
        size_t o = 0;
        while (o + 18 <= buf.size()) {
            if (buf[o] != 0x1f || buf[o + 1] != 0x8b) return false;
            const size_t bsize = (size_t)(buf[o + 16] | buf[o + 17] << 8) + 1;
            if (bsize < 26 || o + bsize > buf.size()) return false;
            const uint32_t isize = rd32(buf.data() + o + bsize - 4);
            const size_t at = raw.size();
            raw.resize(at + isize);
            if (isize) {
                z_stream zs; memset(&zs, 0, sizeof zs);
                if (inflateInit2(&zs, -15) != Z_OK) return false;
                zs.next_in = buf.data() + o + 18; zs.avail_in = (uInt)(bsize - 26);
                zs.next_out = raw.data() + at; zs.avail_out = isize;
                const int zr = inflate(&zs, Z_FINISH);
                inflateEnd(&zs);
                if (zr != Z_STREAM_END || zs.avail_out != 0) return false;
            }
            o += bsize;
        }
        buf.swap(raw);
        if (buf.size() < 8) return false;
    }
    const uint8_t* p = buf.data() + 4; const uint8_t* e = buf.data() + buf.size();
    auto need = [&](size_t n) { return (size_t)(e - p) >= n; };
    if (memcmp(buf.data(), "CSI\1", 4) == 0) {                  // hts.c:1580-1594
        if (!need(12)) return false;
        min_shift = rdi32(p); n_lvls = rdi32(p + 4); const int32_t l_aux = rdi32(p + 8); p += 12;
        if (min_shift < 0 || min_shift > 30 || n_lvls < 1 || n_lvls > 9 || l_aux < 0 || !need((size_t)l_aux)) return false;
        p += l_aux;
        is_csi = true;
    } else if (memcmp(buf.data(), "BAI\1", 4) == 0) {
        min_shift = 14; n_lvls = 5; is_csi = false;
    } else {
        return false;
    }
    if (!need(4)) return false;
    int32_t n_ref = rdi32(p); p += 4;
    if (n_ref < 0) return false;
    refs.assign((size_t)n_ref, Ref());
    for (int32_t i = 0; i < n_ref; ++i) {
        Ref& r = refs[i];
        if (!need(4)) return false;
        int32_t n_bin = rdi32(p); p += 4;
        if (n_bin < 0) return false;
        r.bins.reserve((size_t)n_bin);
        for (int32_t j = 0; j < n_bin; ++j) {
            if (!need(is_csi ? 16 : 8)) return false;
            Bin b; b.bin = rd32(p); p += 4;
            b.loff = 0;
            if (is_csi) { b.loff = rd64(p); p += 8; }              // hts_idx_load_core, hts.c:1535-1538
            int32_t n_chunk = rdi32(p); p += 4;
            if (n_chunk < 0 || !need((size_t)n_chunk * 16)) return false;
            b.chunks.resize((size_t)n_chunk);
            for (int32_t k = 0; k < n_chunk; ++k) { b.chunks[k].beg = rd64(p); b.chunks[k].end = rd64(p + 8); p += 16; }
            r.bins.push_back(std::move(b));
        }
        std::sort(r.bins.begin(), r.bins.end(), [](const Bin& a, const Bin& b) { return a.bin < b.bin; });
        if (is_csi) continue;                                     // no linear index in a CSI
        if (!need(4)) return false;
        int32_t n_intv = rdi32(p); p += 4;
        if (n_intv < 0 || !need((size_t)n_intv * 8)) return false;
        r.ioffset.resize((size_t)n_intv);
        for (int32_t k = 0; k < n_intv; ++k) { r.ioffset[k] = rd64(p); p += 8; }
        for (int32_t k = 1; k < n_intv; ++k) if (r.ioffset[k] == 0) r.ioffset[k] = r.ioffset[k - 1];   // hts.c:1559-1560
        // update_loff (hts.c:1193-1222): loff of a bin = linear-index entry of its first 16 kb window
        for (Bin& b : r.bins) {
            if (b.bin >= 37449u) continue;
            int l = 0;
            for (uint32_t t = b.bin; t; t = (t - 1) >> 3) ++l;
            const uint32_t first = ((1u << (3 * l)) - 1u) / 7u;
            const int64_t bot = (int64_t)(b.bin - first) << (3 * (5 - l));
            b.loff = bot < n_intv ? r.ioffset[(size_t)bot] : 0;
        }
    }
    n_no_coor = need(8) ? rd64(p) : 0;
    return true;
}

bool BaiIndex::load_for_bam(const std::string& bam, BaiIndex* out, bool* csi_present) {
    auto exists = [](const std::string& f) { struct stat st; return stat(f.c_str(), &st) == 0; };
    std::string stem = bam;
    for (size_t i = bam.size(); i-- > 1;) if (bam[i] == '.') { stem = bam.substr(0, i); break; }
    if (csi_present) *csi_present = false;
    // hts_idx_load (hts.c:2031-2042): ".csi" first (next to the file, then with the extension replaced), then ".bai"
    if (exists(bam + ".csi")) { if (csi_present) *csi_present = true; return out->load(bam + ".csi"); }
    if (exists(stem + ".csi")) { if (csi_present) *csi_present = true; return out->load(stem + ".csi"); }
    if (exists(bam + ".bai")) return out->load(bam + ".bai");
    if (exists(stem + ".bai")) return out->load(stem + ".bai");
    return false;
}

bool BaiIndex::whole_file_start(uint64_t* voff) const {
    uint64_t off0 = UINT64_MAX;
    for (const Ref& r : refs) {
        const Bin* m = r.find(meta_bin());
        if (m && !m->chunks.empty() && off0 > m->chunks[0].beg) off0 = m->chunks[0].beg;
    }
    if (off0 == UINT64_MAX && n_no_coor) off0 = 0;
    if (off0 == UINT64_MAX) return false;
    *voff = off0;
    return true;
}

bool BaiIndex::contig_range(int32_t tid, Chunk64* out) const {
    if (tid < 0 || (size_t)tid >= refs.size()) return false;
    const Bin* m = refs[tid].find(meta_bin());
    if (!m || m->chunks.empty()) return false;
    *out = m->chunks[0];
    return out->end > out->beg;
}

std::vector<Chunk64> BaiIndex::query(int32_t tid, int64_t beg, int64_t end) const {
    std::vector<Chunk64> off;
    if (tid < 0 || (size_t)tid >= refs.size()) return off;
    const Ref& r = refs[tid];
    // min_off: hts.c:1765-1776
    uint64_t min_off = 0;
    {
        uint32_t bin = ((1u << (3 * n_lvls)) - 1u) / 7u + (uint32_t)(beg >> min_shift);      // hts_bin_first(n_lvls)
        const Bin* k = nullptr;
        do {
            k = r.find(bin);
            if (k) break;
            const uint32_t parent = (bin - 1) >> 3, first = (parent << 3) + 1;
            if (bin > first) --bin; else bin = parent;
        } while (bin);
        if (bin == 0) k = r.find(0);
        min_off = k ? k->loff : 0;
    }
    // reg2bins: hts.c:1690-1706
    if (beg < end) {
        int64_t e = end;
        int s = min_shift + 3 * n_lvls;
        if (e >= (1ll << s)) e = 1ll << s;
        --e;
        int t = 0;
        for (int l = 0; l <= n_lvls; s -= 3, t += 1 << (3 * l), ++l)
            for (int64_t b = t + (beg >> s); b <= t + (e >> s); ++b) {
                const Bin* k = r.find((uint32_t)b);
                if (!k) continue;
                for (const Chunk64& c : k->chunks) if (c.end > min_off) off.push_back(c);
            }
    }
    if (off.empty()) return off;
    std::sort(off.begin(), off.end(), [](const Chunk64& a, const Chunk64& b) { return a.beg < b.beg; });
    size_t l = 0;
    for (size_t i = 1; i < off.size(); ++i) if (off[l].end < off[i].end) off[++l] = off[i];
    off.resize(l + 1);
    for (size_t i = 1; i < off.size(); ++i) if (off[i - 1].end >= off[i].beg) off[i - 1].end = off[i].beg;
    l = 0;
    for (size_t i = 1; i < off.size(); ++i) {
        if (off[l].end >> 16 == off[i].beg >> 16) off[l].end = off[i].end; else off[++l] = off[i];
    }
    off.resize(l + 1);
    return off;
}

// ---- region string -------------------------------------------------------------------------------
namespace {
// hts_parse_decimal with HTS_PARSE_THOUSANDS_SEP (hts.c:1833-1875)
long long parse_decimal(const char* s, const char** end) {
    long long n = 0; int decimals = 0, e = 0; char sign = '+';
    while (*s == ' ' || *s == '\t' || *s == '\n' || *s == '\v' || *s == '\f' || *s == '\r') ++s;
    if (*s == '+' || *s == '-') sign = *s++;
    while (*s) { if (*s >= '0' && *s <= '9') n = 10 * n + (*s++ - '0'); else if (*s == ',') ++s; else break; }
    if (*s == '.') { ++s; while (*s >= '0' && *s <= '9') { ++decimals; n = 10 * n + (*s++ - '0'); } }
    if (*s == 'E' || *s == 'e') { char* t; e = (int)strtol(s + 1, &t, 10); s = t; }
    e -= decimals;
    while (e > 0) { n *= 10; --e; }
    while (e < 0) { n /= 10; ++e; }
    if (end) *end = s;
    return sign == '+' ? n : -n;
}
}  // namespace

bool parse_region(const BamFile& bam, const std::string& region, IterSpec* out) {
    if (region == ".") { out->kind = IterSpec::WholeFile; return true; }
    if (region == "*") { out->kind = IterSpec::NoCoor; return true; }
    const char* s = region.c_str();
    const char* colon = strrchr(s, ':');
    int beg = 0, end = INT_MAX; bool parsed = true;
    if (!colon) colon = s + region.size();
    else {
        const char* hy;
        beg = (int)(parse_decimal(colon + 1, &hy) - 1);
        if (beg < 0) beg = 0;
        if (*hy == '\0') end = INT_MAX;
        else if (*hy == '-') end = (int)parse_decimal(hy + 1, nullptr);
        else parsed = false;
        if (parsed && beg >= end) parsed = false;
    }
    int32_t tid;
    if (parsed) tid = bam.name2id(std::string(s, (size_t)(colon - s)));
    else { tid = bam.name2id(region); beg = 0; end = INT_MAX; }
    if (tid < 0) return false;
    out->kind = IterSpec::Region; out->tid = tid; out->beg = beg; out->end = end;
    return true;
}

// ---- driver --------------------------------------------------------------------------------------
bool feed_alignments(const BamFile& bam, const BaiIndex& idx, const IterSpec& spec, const FeederOptions& opt,
                     BatchSink* sink, FeederStats* stats, std::string* err) {
    BatchWriter w(sink, opt, (int32_t)bam.header().names.size());
    switch (spec.kind) {
    case IterSpec::WholeFile: {
        uint64_t off0;
        if (!idx.whole_file_start(&off0)) { if (err) *err = "no alignments indexed"; return false; }
        // curr_off == 0 means "do not seek": keep reading right after the header (hts.c:1929-1932)
        if (off0 == 0) off0 = bam.header().first_record_voffset;
        feed_stream(bam, {Chunk64{off0, UINT64_MAX}}, opt, w, stats);
        break; }
    case IterSpec::NoCoor: {
        uint64_t off0 = UINT64_MAX;
        if (!idx.refs.empty()) {
            const BaiIndex::Bin* m = idx.refs.back().find(idx.meta_bin());
            if (m && !m->chunks.empty()) off0 = m->chunks[0].end;
        }
        if (off0 == UINT64_MAX && idx.n_no_coor) off0 = 0;
        if (off0 == UINT64_MAX) { if (err) *err = "no unplaced alignments indexed"; return false; }
        if (off0 == 0) off0 = bam.header().first_record_voffset;
        feed_stream(bam, {Chunk64{off0, UINT64_MAX}}, opt, w, stats);
        break; }
    case IterSpec::Region: {
        if (spec.end < spec.beg || (size_t)spec.tid >= idx.refs.size()) { if (err) *err = "region outside the index"; return false; }
        feed_region(bam, idx.query(spec.tid, spec.beg, spec.end), spec, w, stats);
        break; }
    case IterSpec::Contigs: {
        feed_stream(bam, coalesced_contig_ranges(idx, spec.contigs), opt, w, stats);
        break; }
    }
    w.finish();
    if (stats) { stats->reads += w.total_reads(); stats->cigar_ops += w.total_ops(); }
    return true;
}

uint64_t scan_bgzf_blocks(const BamFile& bam, uint64_t coff, uint64_t end_coff, size_t max_blocks, uint64_t max_comp_bytes,
                          std::vector<BgzfBlockInfo>* out, bool* stop) {
    *stop = false;
    uint64_t taken = 0;
    while (out->size() < max_blocks && taken < max_comp_bytes) {
        if (coff >= end_coff) { *stop = true; break; }
        BlockDesc d;
        int rc = peek_block(bam.data(), bam.size(), coff, &d);
        if (rc != 0 || d.isize == 0 || d.isize > 0x10000) { *stop = true; break; }
        out->push_back(BgzfBlockInfo{d.coff, d.csize, d.isize});
        coff = d.coff + d.csize; taken += d.csize;
    }
    return coff;
}

uint64_t scan_bgzf_blocks_mem(const uint8_t* buf, size_t n, uint64_t base_coff, uint64_t end_coff,
                              std::vector<BgzfBlockInfo>* out, bool* stop, bool* partial, bool* untrusted) {
    *stop = false; *partial = false;
    if (untrusted) *untrusted = false;
    size_t o = 0;
    for (;;) {
        if (base_coff + o >= end_coff) { *stop = true; break; }
        if (o + 18 > n) { *partial = true; break; }
        const uint8_t* h = buf + o;
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || rd16(h + 10) != 6 || h[12] != 'B' || h[13] != 'C' ||
            rd16(h + 14) != 2) { *stop = true; break; }
        const uint32_t bsize = (uint32_t)rd16(h + 16) + 1u;
        if (bsize < 26) { *stop = true; break; }
        if (o + bsize > n) { *partial = true; break; }
        const uint32_t isize = rd32(h + bsize - 4);
        if (isize == 0 || isize > 0x10000) {
            // htslib never reads ISIZE (a block is as long as zlib says, bgzf.c:315).  The 28-byte empty block really is the end
            // of the stream; any other block claiming 0 or > 64 KB is left to the host feeder, which inflates it to find out.
            if (untrusted && !(isize == 0 && bsize == 28)) *untrusted = true;
            *stop = true; break;
        }
        out->push_back(BgzfBlockInfo{base_coff + o, bsize, isize});
        o += bsize;
    }
    return base_coff + o;
}

// Contigs -> shards.  Each shard is a run of CONSECUTIVE contigs (so that a rank streams one contiguous byte range of
// the coordinate-sorted file: a shard made of scattered contigs is a dozen small ranges, each with its own partly filled
// inflate launches — measured 3.5x slower end to end), balanced on compressed bytes from the BAI pseudo-bin (hts.c:1092):
// the smallest achievable maximum load over all contiguous partitions, then spare ranks split the heaviest runs.
std::vector<int32_t> plan_contig_shards(const BamFile& bam, const BaiIndex& idx, int world) {
    const size_t n = bam.header().names.size();
    std::vector<int32_t> assign(n, 0);
    if (world <= 1 || n == 0) return assign;
    std::vector<uint64_t> w(n, 0);
    uint64_t total = 0, wmax = 0;
    for (size_t t = 0; t < n; ++t) {
        Chunk64 c;
        if (idx.contig_range((int32_t)t, &c)) w[t] = (c.end >> 16) - (c.beg >> 16) + 1;
        total += w[t]; wmax = std::max(wmax, w[t]);
    }
    auto groups_needed = [&](uint64_t cap) {
        int g = 1; uint64_t cur = 0;
        for (size_t t = 0; t < n; ++t) { if (cur + w[t] > cap && cur) { ++g; cur = 0; } cur += w[t]; }
        return g;
    };
    uint64_t lo = wmax, hi = std::max(total, wmax);
    while (lo < hi) { const uint64_t mid = lo + (hi - lo) / 2; if (groups_needed(mid) <= world) hi = mid; else lo = mid + 1; }
    // cut points under the optimal cap
    std::vector<size_t> start{0};                      // first contig of every run
    { uint64_t cur = 0; for (size_t t = 0; t < n; ++t) { if (cur + w[t] > lo && cur) { start.push_back(t); cur = 0; } cur += w[t]; } }
    // spare ranks: split the heaviest run that still holds two contigs with reads, at its most balanced point
    while ((int)start.size() < world) {
        size_t best = SIZE_MAX, best_cut = 0; uint64_t best_load = 0;
        for (size_t g = 0; g < start.size(); ++g) {
            const size_t a0 = start[g], a1 = g + 1 < start.size() ? start[g + 1] : n;
            uint64_t load = 0; size_t nz = 0;
            for (size_t t = a0; t < a1; ++t) { load += w[t]; nz += w[t] != 0; }
            if (nz < 2 || load <= best_load) continue;
            uint64_t left = 0, best_diff = UINT64_MAX; size_t cut = 0;
            for (size_t t = a0; t + 1 < a1; ++t) {
                left += w[t];
                if (left == 0 || left == load) continue;
                const uint64_t diff = left > load - left ? left - (load - left) : (load - left) - left;
                if (diff < best_diff) { best_diff = diff; cut = t + 1; }
            }
            if (cut) { best = g; best_cut = cut; best_load = load; }
        }
        if (best == SIZE_MAX) break;
        start.insert(start.begin() + (long)best + 1, best_cut);
    }
    for (size_t g = 0; g < start.size(); ++g) {
        const size_t a1 = g + 1 < start.size() ? start[g + 1] : n;
        for (size_t t = start[g]; t < a1; ++t) assign[t] = (int32_t)g;
    }
    return assign;
}

// Byte ranges of a set of contigs; contigs that follow each other in the file (nothing but read-less contigs in
// between) become ONE range.
std::vector<Chunk64> coalesced_contig_ranges(const BaiIndex& idx, const std::vector<int32_t>& contigs) {
    std::vector<int32_t> tids(contigs);
    std::sort(tids.begin(), tids.end());
    std::vector<Chunk64> ranges;
    int32_t last = -2;
    for (int32_t tid : tids) {
        Chunk64 c;
        if (!idx.contig_range(tid, &c)) continue;
        bool adjacent = !ranges.empty() && c.beg >= ranges.back().end;
        for (int32_t t = last + 1; adjacent && t < tid; ++t) { Chunk64 x; if (idx.contig_range(t, &x)) adjacent = false; }
        if (adjacent) ranges.back().end = c.end; else ranges.push_back(c);
        last = tid;
    }
    return ranges;
}

}  // namespace rtjx
