// regtools_b200/csrc/stage_pipe.h — read-ahead of the compressed BAM into pinned memory (device feeder, file mode).
//
// Replaces, for the device feeder, what htslib's bgzf_read_block + hread do for the reference (bgzf.c:525-546: one block at a
// time, on the thread that also inflates and parses): the bytes leave the page cache on a pool of threads, ahead of the thread
// that walks the BGZF headers and enqueues device work.  No CUDA in this header — the caller supplies the "buffer is free
// again" wait — so the class is exercised on the CPU (tools/stage_pipe_check.cc, also under ThreadSanitizer).
#ifndef RTJX_STAGE_PIPE_H_
#define RTJX_STAGE_PIPE_H_

#include <stdint.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace rtjx {

// Read-ahead of the compressed file into a ring of pinned buffers, by a pool of threads that lives as long as the run: the
// run's own thread walks BGZF headers and enqueues device work while the next windows are being copied out of the page
// cache.  get(off) hands out the window [off, off + window) (short at the end of the file); windows asked for in sequence
// (`stride` apart, below `seq_end`) have been requested `DEPTH` ahead, anything else is read on demand.  The buffer a window
// lands in is reused NBUF windows later; `wait_free(b)` returns once the caller's last H2D copy out of buffer b has finished (an event wait).
class StagePipe {
public:
    static constexpr int NBUF = 6, DEPTH = 3;
    StagePipe(int fd, uint64_t file_size, uint8_t* const* bufs, std::function<void(int)> wait_free, size_t window, size_t stride, int threads)
        : fd_(fd), file_size_(file_size), bufs_(bufs), wait_free_(std::move(wait_free)), window_(window), stride_(stride), n_threads_(std::max(1, threads)) {
        for (int t = 0; t < n_threads_; ++t) pool_.emplace_back([this, t] { worker(t); });
    }
    ~StagePipe() {
        { std::unique_lock<std::mutex> lk(mu_); stop_ = true; }
        cv_work_.notify_all();
        for (auto& th : pool_) th.join();
    }
    StagePipe(const StagePipe&) = delete;
    StagePipe& operator=(const StagePipe&) = delete;

    // nullptr = short read (I/O error)
    const uint8_t* get(uint64_t off, uint64_t seq_end, size_t* got, int* buf) {
        if (head_ < published_ && jobs_[head_ % NBUF].off != off) {     // not the window that was read ahead: let the pool finish, start over
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [&] { return completed_ == published_; });
            head_ = published_;
        }
        if (head_ == published_) { next_off_ = off; publish(next_off_); next_off_ += stride_; }
        while (published_ - head_ < (uint64_t)DEPTH + 1 && next_off_ < seq_end && next_off_ < file_size_) { publish(next_off_); next_off_ += stride_; }
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_done_.wait(lk, [&] { return completed_ > head_; });
        }
        Job& j = jobs_[head_ % NBUF];
        *buf = (int)(head_ % NBUF);
        *got = j.len;
        ++head_;
        return j.bad.load() ? nullptr : bufs_[*buf];
    }

private:
    struct Job { uint64_t off = 0; size_t len = 0; std::atomic<int> left{0}; std::atomic<bool> bad{false}; };
    void publish(uint64_t off) {
        const int b = (int)(published_ % NBUF);
        wait_free_(b);                                     // the H2D copy that last read this buffer (enqueued >= 2 windows ago)
        Job& j = jobs_[b];
        j.off = off; j.len = (size_t)std::min<uint64_t>(window_, file_size_ - off); j.left.store(n_threads_); j.bad.store(false);
        { std::unique_lock<std::mutex> lk(mu_); ++published_; }
        cv_work_.notify_all();
    }
    void worker(int t) {
        uint64_t seq = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_work_.wait(lk, [&] { return stop_ || published_ > seq; });
                if (published_ <= seq) return;              // stop requested and nothing left to read
            }
            Job& j = jobs_[seq % NBUF];
            const size_t per = ((j.len + (size_t)n_threads_ - 1) / (size_t)n_threads_ + 4095) & ~(size_t)4095;
            const size_t o = (size_t)t * per;
            if (o < j.len) {
                const size_t len = std::min(per, j.len - o);
                size_t g = 0;
                while (g < len) {
                    const ssize_t r = pread(fd_, bufs_[seq % NBUF] + o + g, len - g, (off_t)(j.off + o + g));
                    if (r <= 0) break;
                    g += (size_t)r;
                }
                if (g < len) j.bad.store(true);
            }
            if (j.left.fetch_sub(1) == 1) {
                { std::unique_lock<std::mutex> lk(mu_); ++completed_; }
                cv_done_.notify_all();
            }
            ++seq;
        }
    }
    const int fd_;
    const uint64_t file_size_;
    uint8_t* const* bufs_;
    std::function<void(int)> wait_free_;
    const size_t window_, stride_;
    const int n_threads_;
    Job jobs_[NBUF];
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    uint64_t published_ = 0, completed_ = 0;               // guarded by mu_ (published_ is written by the caller's thread only)
    uint64_t head_ = 0, next_off_ = 0;                     // caller's thread only
    bool stop_ = false;
    std::vector<std::thread> pool_;
};

}  // namespace rtjx

#endif  // RTJX_STAGE_PIPE_H_
