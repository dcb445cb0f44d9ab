// tools/feeder_tsan.cc — developer check (SURVEY §5 race-detection row): the threaded host feeder (bam_feeder.cc: BGZF inflate
// workers + ordered record split) under ThreadSanitizer.  Built and run by `make -C tools tsan` style one-liner:
//   g++ -O1 -g -fsanitize=thread -std=c++17 -pthread -I regtools_b200/csrc tools/feeder_tsan.cc regtools_b200/csrc/bam_feeder.cc -lz -o build/feeder_tsan
//   build/feeder_tsan <bam> [region] [threads]
// Streams the whole file (or a region) through feed_alignments with a sink that sums what it sees; prints the totals.
#include "bam_feeder.h"

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

using namespace rtjx;

struct SumSink : BatchSink {
    std::vector<std::unique_ptr<HostBatch>> pool;
    std::vector<std::vector<int32_t>> i32; std::vector<std::vector<uint32_t>> u32;
    size_t next = 0;
    unsigned long long reads = 0, ops = 0, nops = 0, sum = 0;
    SumSink() {
        for (int i = 0; i < 3; ++i) {
            std::unique_ptr<HostBatch> b(new HostBatch());
            const uint32_t R = 1u << 14, O = 1u << 17;
            i32.emplace_back(R); b->tid = i32.back().data();
            i32.emplace_back(R); b->pos = i32.back().data();
            u32.emplace_back(R); b->meta = u32.back().data();
            u32.emplace_back(R + 1); b->cig_off = u32.back().data();
            u32.emplace_back(O); b->cigar = u32.back().data();
            b->cap_reads = R; b->cap_ops = O;
            pool.push_back(std::move(b));
        }
    }
    HostBatch* acquire() override { HostBatch* b = pool[next].get(); next = (next + 1) % pool.size(); return b; }
    void submit(HostBatch* b) override {
        reads += b->n_reads; ops += b->n_ops; nops += b->n_junction_ops;
        for (uint32_t i = 0; i < b->n_reads; ++i) sum += (unsigned)b->pos[i] ^ b->meta[i];
    }
};

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: feeder_tsan in.bam [region] [threads]\n"); return 2; }
    BamFile bam; std::string err;
    if (!bam.open(argv[1], &err)) { fprintf(stderr, "open: %s\n", err.c_str()); return 1; }
    BaiIndex idx; bool csi = false;
    if (!BaiIndex::load_for_bam(argv[1], &idx, &csi)) { fprintf(stderr, "no index\n"); return 1; }
    IterSpec spec;
    if (!parse_region(bam, argc > 2 ? argv[2] : ".", &spec)) { fprintf(stderr, "bad region\n"); return 1; }
    FeederOptions fo; fo.n_threads = argc > 3 ? atoi(argv[3]) : 4;
    SumSink sink; FeederStats fs;
    if (!feed_alignments(bam, idx, spec, fo, &sink, &fs, &err)) { fprintf(stderr, "feed: %s\n", err.c_str()); return 1; }
    printf("reads %llu ops %llu junction_ops %llu checksum %llu blocks %llu\n", sink.reads, sink.ops, sink.nops, sink.sum, (unsigned long long)fs.bgzf_blocks);
    return 0;
}
