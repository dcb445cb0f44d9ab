// regtools_b200/csrc/rtjx_api.cc — the C ABI (include/rtjx.h): thin, exception-free forwarding to Engine.
#include "engine.h"

#include <cstring>
#include <new>

using rtjx::Engine;

struct rtjx_handle { Engine* e; };

static thread_local char g_create_err[256];

extern "C" {

void rtjx_params_default(rtjx_params* p) {
    if (!p) return;
    memset(p, 0, sizeof *p);
    p->struct_size = (uint32_t)sizeof *p;
    p->region = ".";
    p->strand_tag = "XS";
    p->strandness = 0;
    p->min_anchor = 8;          // junctions_extractor.h:185-187
    p->min_intron = 70;
    p->max_intron = 500000;
    p->device = 0;
    p->shard_world = 1;
}

int rtjx_create(const rtjx_params* p, rtjx_t** out) {
    if (!p || !out) return RTJX_E_ARG;
    *out = nullptr;
    if (p->struct_size != sizeof(rtjx_params)) { snprintf(g_create_err, sizeof g_create_err, "rtjx_params.struct_size mismatch"); return RTJX_E_ARG; }
    if (p->barcode_out && p->shard_world > 1) { snprintf(g_create_err, sizeof g_create_err, "-b barcodes are not exchanged between contig shards"); return RTJX_E_UNSUPPORTED; }
    if (p->strandness < 0 || p->strandness > 3) { snprintf(g_create_err, sizeof g_create_err, "strandness must be 0..3"); return RTJX_E_ARG; }
    if (p->shard_world > 1 && (p->shard_rank < 0 || p->shard_rank >= p->shard_world)) { snprintf(g_create_err, sizeof g_create_err, "shard_rank out of range"); return RTJX_E_ARG; }
    rtjx_handle* h = new (std::nothrow) rtjx_handle;
    if (!h) return RTJX_E_NOMEM;
    try { h->e = new Engine(*p); } catch (...) { delete h; return RTJX_E_NOMEM; }
    *out = h;
    return RTJX_OK;
}

void rtjx_destroy(rtjx_t* h) {
    if (!h) return;
    delete h->e;
    delete h;
}

#define GUARD(h, expr)                                                     \
    if (!(h)) return RTJX_E_ARG;                                           \
    try { return (expr); }                                                 \
    catch (const std::bad_alloc&) { return (h)->e->fail(RTJX_E_NOMEM, "out of host memory"); } \
    catch (const std::exception& ex) { return (h)->e->fail(RTJX_E_STATE, ex.what()); }          \
    catch (...) { return (h)->e->fail(RTJX_E_STATE, "unknown internal error"); }

int rtjx_run(rtjx_t* h) { GUARD(h, h->e->run()) }
int rtjx_run_regions(rtjx_t* h, const char* const* regions, size_t n) { GUARD(h, h->e->run_regions(regions, n)) }
int64_t rtjx_region_count(rtjx_t* h, size_t region) { GUARD(h, h->e->region_count(region)) }
int64_t rtjx_region_get(rtjx_t* h, size_t region, rtjx_junction* out, size_t cap) { GUARD(h, h->e->region_get(region, out, cap)) }
int rtjx_unique_junctions(rtjx_t* h, const uint32_t* win_start, const uint32_t* win_end, size_t n_regions) { GUARD(h, h->e->unique_build(win_start, win_end, n_regions)) }
int64_t rtjx_unique_count(rtjx_t* h) { GUARD(h, h->e->unique_count()) }
int64_t rtjx_unique_get(rtjx_t* h, rtjx_junction* out, uint32_t* first_region, size_t cap) { GUARD(h, h->e->unique_get(out, first_region, cap)) }
int64_t rtjx_unique_regions(rtjx_t* h, size_t i, uint32_t* out, size_t cap) { GUARD(h, h->e->unique_regions(i, out, cap)) }

int rtjx_scan_batch(rtjx_t* h, const rtjx_batch* b, int location, void* stream) {
    if (!h) return RTJX_E_ARG;
    if (!b) return h->e->fail(RTJX_E_ARG, "null batch");
    GUARD(h, h->e->scan_batch(*b, location, static_cast<cudaStream_t>(stream)))
}

int rtjx_add(rtjx_t* h, const rtjx_candidate* c, size_t n) { GUARD(h, h->e->add(c, n)) }
int rtjx_finalize(rtjx_t* h, void* stream) { GUARD(h, h->e->finalize(static_cast<cudaStream_t>(stream))) }
int64_t rtjx_count(rtjx_t* h) { GUARD(h, h->e->count()) }
int64_t rtjx_get(rtjx_t* h, rtjx_junction* out, size_t cap) { GUARD(h, h->e->get(out, cap)) }
int rtjx_write_bed12(rtjx_t* h, int fd) { GUARD(h, h->e->write_bed12(fd)) }
int64_t rtjx_intern_barcode(rtjx_t* h, const char* barcode) { GUARD(h, h->e->intern_barcode(barcode)) }
int rtjx_write_barcodes(rtjx_t* h, int fd) { GUARD(h, h->e->write_barcodes(fd)) }
int rtjx_barcode_stats(rtjx_t* h, uint64_t* n_barcodes, uint64_t* n_missing) { GUARD(h, h->e->barcode_stats(n_barcodes, n_missing)) }
const char* rtjx_barcode_name(rtjx_t* h, uint32_t id) { return h ? h->e->barcode_name(id) : nullptr; }
int64_t rtjx_load_barcodes(rtjx_t* h, uint32_t* ids, size_t cap) { GUARD(h, h->e->load_barcodes(ids, cap)) }
int rtjx_import(rtjx_t* h, const rtjx_junction* j, size_t n) { GUARD(h, h->e->import(j, n)) }
int rtjx_clear(rtjx_t* h) { GUARD(h, h->e->clear()) }

int rtjx_load_batch(rtjx_t* h, uint64_t* n_reads, uint64_t* n_ops, int32_t* tid, int32_t* pos, uint32_t* meta,
                    uint32_t* cig_off, uint32_t* cigar) {
    GUARD(h, h->e->load_batch(n_reads, n_ops, tid, pos, meta, cig_off, cigar))
}

int rtjx_inflate_file(rtjx_t* h, uint64_t max_blocks, void* out, uint64_t cap, uint64_t* out_len) {
    GUARD(h, h->e->inflate_file(max_blocks, out, cap, out_len))
}

int rtjx_comm_unique_id(void* id) {
    std::string err;
    if (!id) return RTJX_E_ARG;
    int rc = rtjx::comm_unique_id(id, &err);
    if (rc) snprintf(g_create_err, sizeof g_create_err, "%s", err.c_str());
    return rc;
}
int rtjx_comm_init(rtjx_t* h, const void* id, int rank, int world) {
    if (!h || !id) return RTJX_E_ARG;
    std::string err;
    int rc = rtjx::comm_init(id, rank, world, h->e->device(), &err);
    if (rc) return h->e->fail(rc, err);
    return RTJX_OK;
}
void rtjx_comm_destroy(void) { rtjx::comm_destroy(); }
int rtjx_gather(rtjx_t* h, int root) { GUARD(h, h->e->gather(root)) }

int rtjx_stage_bam(rtjx_t* h) {
    GUARD(h, h->e->stage_file())
}

const char* rtjx_contig(rtjx_t* h, int32_t tid) { return h ? h->e->contig(tid) : ""; }
int32_t rtjx_n_contigs(rtjx_t* h) { return h ? h->e->n_contigs() : 0; }
int32_t rtjx_intern_contig(rtjx_t* h, const char* name) { return h ? h->e->intern_contig(name) : -1; }

int32_t rtjx_plan_shards(const char* bam, int32_t world, int32_t* assign, size_t cap) {
    if (!bam || world < 1) return RTJX_E_ARG;
    try {
        rtjx::BamFile f; std::string err;
        if (!f.open(bam, &err)) return RTJX_E_OPEN_BAM;
        rtjx::BaiIndex idx; bool csi = false;
        if (!rtjx::BaiIndex::load_for_bam(bam, &idx, &csi)) return RTJX_E_OPEN_INDEX;
        std::vector<int32_t> a = rtjx::plan_contig_shards(f, idx, world);
        for (size_t i = 0; i < a.size() && i < cap && assign; ++i) assign[i] = a[i];
        return (int32_t)a.size();
    } catch (...) { return RTJX_E_NOMEM; }
}

int rtjx_get_stats(rtjx_t* h, rtjx_stats* out) {
    if (!h || !out) return RTJX_E_ARG;
    h->e->get_stats(out);
    return RTJX_OK;
}
void rtjx_reset_stats(rtjx_t* h) { if (h) h->e->reset_stats(); }

const char* rtjx_last_error(const rtjx_t* h) { return h ? h->e->last_error() : g_create_err; }

const char* rtjx_strerror(int status) {
    switch (status) {
    case RTJX_OK: return "ok";
    case RTJX_E_ARG: return "invalid argument";
    case RTJX_E_OPEN_BAM: return "Unable to open BAM/SAM file.\n\n";
    case RTJX_E_OPEN_INDEX: return "Unable to open BAM/SAM index. Make sure alignments are indexed\n\n";
    case RTJX_E_REGION: return "Unable to iterate to region within BAM.\n\n";
    case RTJX_E_CUDA: return "CUDA device unavailable or CUDA call failed";
    case RTJX_E_UNSUPPORTED: return "not supported by the B200 path";
    case RTJX_E_NOMEM: return "out of memory";
    case RTJX_E_STATE: return "internal state error";
    case RTJX_E_IO: return "I/O error";
    default: return "unknown status";
    }
}

const char* rtjx_version(void) { return "regtools-b200 0.1 (junctions extract; reference regtools 1.0.0)"; }

}  // extern "C"
