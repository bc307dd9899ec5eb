// dh_hostenc.hpp — host side of the compressed host->device path of dh_predict_batch: a zero-run
// encoder that writes frames in the Biwi depth-file format (src/db_reader/biwi.rs:81-103, the
// format biwi_decode_kernel expands on the GPU) and the small worker pool that runs it.
#pragma once

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace dh {

// Upper bound of the encoded size of one w x h frame: header + every pixel + one run header per
// alternation of 16-pixel groups (a run is at least one group, except at the frame's tail).
size_t rle_frame_bound(uint32_t w, uint32_t h);

// Encodes one frame.  Runs are found at a granularity of 16 pixels (32 bytes): a group with any
// non-zero pixel belongs to a "full" run and is copied verbatim (isolated zero pixels inside it
// travel as literal zeros, which the format allows), groups of 16 zeros extend an "empty" run.
// The stream read_depth (biwi.rs:81-103) produces from the result is exactly `src`.  Returns the
// number of bytes written (a multiple of 4; dst needs rle_frame_bound() bytes).
size_t rle_encode_frame(const uint16_t* src, uint32_t w, uint32_t h, uint8_t* dst);

// Fraction of 16-pixel groups with a non-zero pixel among every `step`-th group of a frame
// (a cheap estimate of what the encoder would keep).
double rle_sample_density(const uint16_t* src, size_t npx, size_t step);

// A fixed set of worker threads executing index ranges: run(n, fn) calls fn(i) for i in [0, n)
// on the workers (dynamic hand-out through an atomic counter) and returns a ticket; wait(ticket)
// blocks until that job is done.  Jobs are executed in submission order; several may be queued.
class WorkerPool {
public:
    explicit WorkerPool(unsigned n_threads);
    ~WorkerPool();
    WorkerPool(const WorkerPool&) = delete;
    WorkerPool& operator=(const WorkerPool&) = delete;
    unsigned size() const { return (unsigned)threads_.size(); }
    uint64_t run(uint32_t n, std::function<void(uint32_t)> fn);
    void wait(uint64_t ticket);
    bool done(uint64_t ticket);                          // non-blocking
    bool wait_for(uint64_t ticket, unsigned micros);     // true if the job finished within the time

private:
    struct Job {
        uint64_t ticket = 0;
        uint32_t n = 0;
        std::function<void(uint32_t)> fn;
        std::atomic<uint32_t> next{0};
        std::atomic<uint32_t> done{0};
    };
    void worker();
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_work_, cv_done_;
    std::vector<Job*> queue_;   // jobs not yet fully handed out, oldest first
    uint64_t next_ticket_ = 1, finished_upto_ = 0;  // every ticket <= finished_upto_ is complete
    std::vector<uint64_t> finished_;               // completed tickets above finished_upto_
    bool stop_ = false;
};

// Threads the pool should use by default: the cores this process may run on (sched_getaffinity),
// capped; DH_ENCODE_THREADS overrides.
unsigned default_encode_threads();

}  // namespace dh
