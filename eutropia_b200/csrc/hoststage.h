// Host side of the column transfers when the caller's matrix is PAGEABLE memory (what R hands over:
// REALSXP data, src/RcppExports.cpp:30-43).  cudaMemcpyAsync from / to pageable memory is staged by
// the driver through a small internal buffer, synchronously and at a fraction of the link rate, so the
// one-shot pipeline (H2D | substeps | D2H) would collapse into a sequence.  Here the library stages
// itself: a ring of pinned slots per direction and device, filled / drained by a few copier threads
// with non-temporal stores, the DMA of one slot overlapping the memcpy of the next.
//   up:   caller thread: [wait slot] -> parallel memcpy(src -> slot) -> cudaMemcpyAsync(slot -> device)
//   down: a downloader thread per field takes whole download requests off the caller's hands: it waits
//         (stream ordered) for the substeps of the columns, un-permutes them into the device staging
//         area, then per slot cudaMemcpyAsync(device -> slot) -> [event] -> parallel memcpy(slot -> dst).
//         The caller thread never blocks on a download, so the substep launches of the next chunk and the
//         fills of the one after are not held up by results that are not ready yet.
// Pinned (cudaHostAlloc / cudaHostRegister) caller memory keeps the direct path.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

struct eut_field;

namespace eut {

// memcpy with non-temporal stores for the bulk (neither the pinned slot nor the caller's result is read
// again by this core); falls back to memcpy below 4 KB or without AVX2.
void copy_nt(void *dst, const void *src, size_t bytes);

// true when `p` is host memory CUDA can DMA from / to directly (pinned or registered)
bool host_is_pinned(const void *p);

// Threads this process may use for copier pools: the CPUs it is allowed to run on, divided by the number
// of ranks a launcher runs on this host (LOCAL_WORLD_SIZE).
unsigned host_thread_budget();

class CopyPool {
public:
    explicit CopyPool(unsigned n_threads, const std::vector<int> &cpus);
    ~CopyPool();
    // dst[0, bytes) = src[0, bytes), split over the pool's threads (the caller takes a share); blocking
    void copy(void *dst, const void *src, size_t bytes);
    unsigned size() const { return (unsigned)threads_.size() + 1; }

private:
    void worker();
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    struct Job { char *dst; const char *src; size_t bytes; };
    std::deque<Job> jobs_;
    int pending_ = 0;
    bool stop_ = false;
};

// CPUs close to a CUDA device (sysfs local_cpulist of its PCI function) that this process may use;
// empty when unknown.
std::vector<int> device_local_cpus(int device);

struct HostStager {
    static constexpr int kSlots = 4;
    int device = 0;
    size_t slot_bytes = 0;
    char *up_slot[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    char *dn_slot[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t up_ev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t dn_ev[kSlots] = {nullptr, nullptr, nullptr, nullptr};
    bool up_used[kSlots] = {false, false, false, false};
    uint64_t up_count = 0;
    CopyPool *fill = nullptr, *drain = nullptr;
    // downloader
    struct DownJob {
        std::vector<int32_t> cols;   // field columns (0-based); host column j of h_dst receives cols[j]
        double *h_dst = nullptr;
        int64_t dst_ld = 0;
        cudaEvent_t after = nullptr; // recorded by the caller on the stream that produced the columns; owned by the job
        uint64_t ticket = 0;
    };
    struct ::eut_field *field = nullptr;
    int32_t *d_cols = nullptr, *h_cols = nullptr;   // [2][C] column lists of the un-permute launches (device / pinned)
    uint64_t dn_chunks = 0;
    std::thread downloader;
    std::mutex mu;
    std::condition_variable cv, idle_cv;
    std::deque<DownJob> queue;
    uint64_t issued = 0, finished = 0;
    bool stop = false;
    int err = 0;
    std::string err_msg;
    // trace counters (milliseconds since the last take_stats): caller thread blocked in fills / slot waits,
    // downloader blocked in drains / DMA waits
    double ms_fill = 0, ms_up_wait = 0, ms_drain = 0, ms_dn_wait = 0;
    std::string take_stats();

    int init(struct ::eut_field *field, unsigned threads_per_pool);
    ~HostStager();
    // host (pageable) -> device, stream ordered on `stream`; returns after the last slice has been queued
    int upload(double *d_dst, int64_t dst_ld, const double *h_src, int64_t src_ld, int64_t rows, int64_t cols,
               cudaStream_t stream);
    // queue a device -> host (pageable) job; returns at once with a ticket
    uint64_t download(DownJob &&job);
    int wait_ticket(uint64_t ticket);   // blocks until that job has landed in the caller's memory
    int wait_idle();

private:
    void run_downloader();
};

}  // namespace eut
