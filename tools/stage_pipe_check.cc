// tools/stage_pipe_check.cc — CPU check of the device feeder's read-ahead ring (regtools_b200/csrc/stage_pipe.h).
//
//   g++ -O1 -g -std=c++17 [-fsanitize=thread] -pthread tools/stage_pipe_check.cc -o /tmp/stage_pipe_check && /tmp/stage_pipe_check
//
// Writes a file of known bytes, then asks the pipe for windows the way run_device does — sequential runs on a fixed grid,
// jumps to a new range, a range that ends early, the short window at the end of the file — and compares every window with
// the expected bytes.  The "buffer is free" wait is a stub that also checks the ring's promise: a buffer is only handed to
// the pool again after the caller has released the window it held before.
#include "../regtools_b200/csrc/stage_pipe.h"

#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>

#include <random>

static uint8_t expect_byte(uint64_t off) { return (uint8_t)((off * 2654435761u) >> 13 ^ off >> 20); }

int main() {
    const uint64_t FILE_BYTES = (37u << 20) + 12345;
    const size_t STRIDE = 1u << 20, WINDOW = STRIDE + (128u << 10);
    char path[] = "/tmp/stage_pipe_check_XXXXXX";
    int fd = mkstemp(path);
    if (fd < 0) { perror("mkstemp"); return 2; }
    {
        std::vector<uint8_t> chunk(1u << 20);
        for (uint64_t o = 0; o < FILE_BYTES; o += chunk.size()) {
            const size_t n = (size_t)std::min<uint64_t>(chunk.size(), FILE_BYTES - o);
            for (size_t i = 0; i < n; ++i) chunk[i] = expect_byte(o + i);
            if (write(fd, chunk.data(), n) != (ssize_t)n) { perror("write"); return 2; }
        }
    }
    std::vector<std::vector<uint8_t>> store(rtjx::StagePipe::NBUF, std::vector<uint8_t>(WINDOW + 512));
    uint8_t* bufs[rtjx::StagePipe::NBUF];
    for (int i = 0; i < rtjx::StagePipe::NBUF; ++i) bufs[i] = store[(size_t)i].data();
    std::atomic<int> held[rtjx::StagePipe::NBUF];          // 1 while the "caller" still reads the window in that buffer
    for (auto& h : held) h.store(0);
    long bad = 0, windows = 0;
    for (int threads : {1, 3, 8}) {
        rtjx::StagePipe pipe(fd, FILE_BYTES, bufs, [&](int b) { if (held[b].load()) { fprintf(stderr, "buffer %d reused while held\n", b); ++bad; } },
                             WINDOW, STRIDE, threads);
        std::mt19937_64 rng(99 + (uint64_t)threads);
        int last_buf = -1;
        auto take = [&](uint64_t off, uint64_t seq_end) {
            if (last_buf >= 0) held[last_buf].store(0);     // the H2D copy out of the previous window "finished"
            size_t got = 0; int b = -1;
            const uint8_t* w = pipe.get(off, seq_end, &got, &b);
            ++windows;
            if (!w) { fprintf(stderr, "short read at %llu\n", (unsigned long long)off); ++bad; return; }
            held[b].store(1); last_buf = b;
            const size_t want = (size_t)std::min<uint64_t>(WINDOW, FILE_BYTES - off);
            if (got != want) { fprintf(stderr, "window at %llu: %zu bytes, expected %zu\n", (unsigned long long)off, got, want); ++bad; return; }
            for (size_t i = 0; i < got; i += 97) if (w[i] != expect_byte(off + i)) { fprintf(stderr, "byte %llu differs\n", (unsigned long long)(off + i)); ++bad; return; }
            if (got && w[got - 1] != expect_byte(off + got - 1)) { fprintf(stderr, "last byte of window at %llu differs\n", (unsigned long long)off); ++bad; }
        };
        // whole file, in sequence, from an odd first offset
        for (uint64_t o = 4321; o < FILE_BYTES; o += STRIDE) take(o, FILE_BYTES);
        // ranges: random starts, random lengths, some abandoned before their announced end
        for (int r = 0; r < 40; ++r) {
            const uint64_t beg = rng() % FILE_BYTES, end = std::min<uint64_t>(FILE_BYTES, beg + (rng() % (9u << 20)) + 1);
            const uint64_t stop_after = (rng() & 1) ? end : beg + (end - beg) / 2;
            for (uint64_t o = beg; o < stop_after; o += STRIDE) take(o, end);
        }
        // the same window twice, and a step backwards
        take(5u << 20, FILE_BYTES); take(5u << 20, FILE_BYTES); take(2u << 20, FILE_BYTES);
        if (last_buf >= 0) held[last_buf].store(0);
    }
    close(fd); unlink(path);
    printf("stage_pipe_check: %ld windows, %ld errors\n", windows, bad);
    return bad ? 1 : 0;
}
