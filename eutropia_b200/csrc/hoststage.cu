// Pageable-host staging for the column transfers: see hoststage.h.
#include <sched.h>
#include <pthread.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.h"
#include "hoststage.h"

namespace eut {

bool host_is_pinned(const void *p)
{
    if (getenv("EUT_FORCE_STAGED")) return false;       // treat every host pointer as pageable (tests)
    if (getenv("EUT_NO_STAGING")) return true;          // never stage: pageable memory goes to the driver as it is
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

unsigned host_thread_budget()
{
    cpu_set_t set;
    unsigned n = 0;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) n = (unsigned)CPU_COUNT(&set);
    if (n == 0) n = std::max(1u, std::thread::hardware_concurrency());
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) n = std::max(1u, n / (unsigned)std::max(1, atoi(e)));
    return n;
}

std::vector<int> device_local_cpus(int device)
{
    std::vector<int> cpus;
    char bus[64] = {0};
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) { (void)cudaGetLastError(); return cpus; }
    for (char *c = bus; *c; c++) *c = (char)tolower(*c);
    char path[160];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE *fp = fopen(path, "r");
    if (!fp) return cpus;
    char line[4096] = {0};
    const bool ok = fgets(line, sizeof(line), fp) != nullptr;
    fclose(fp);
    if (!ok) return cpus;
    cpu_set_t allowed;
    const bool have_allowed = sched_getaffinity(0, sizeof(allowed), &allowed) == 0;
    for (char *tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k < 1) continue;
        for (int c = a; c <= b && c < CPU_SETSIZE; c++)
            if (!have_allowed || CPU_ISSET(c, &allowed)) cpus.push_back(c);
    }
    return cpus;
}

static void bind_thread(const std::vector<int> &cpus)
{
    if (cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus) CPU_SET(c, &set);
    (void)pthread_setaffinity_np(pthread_self(), sizeof(set), &set);
}

// ---- CopyPool ---------------------------------------------------------------------------------
CopyPool::CopyPool(unsigned n_threads, const std::vector<int> &cpus)
{
    for (unsigned t = 1; t < std::max(1u, n_threads); t++)
        threads_.emplace_back([this, cpus] { bind_thread(cpus); worker(); });
}

CopyPool::~CopyPool()
{
    { std::lock_guard<std::mutex> l(mu_); stop_ = true; }
    cv_.notify_all();
    for (auto &t : threads_) t.join();
}

void CopyPool::worker()
{
    for (;;) {
        Job j;
        {
            std::unique_lock<std::mutex> l(mu_);
            cv_.wait(l, [this] { return stop_ || !jobs_.empty(); });
            if (jobs_.empty()) return;
            j = jobs_.front();
            jobs_.pop_front();
        }
        copy_nt(j.dst, j.src, j.bytes);
        {
            std::lock_guard<std::mutex> l(mu_);
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
}

void CopyPool::copy(void *dst, const void *src, size_t bytes)
{
    const size_t parts = std::min<size_t>(size(), std::max<size_t>(1, bytes / (256 << 10)));
    if (parts <= 1) { copy_nt(dst, src, bytes); return; }
    const size_t per = ((bytes / parts) + 4095) & ~(size_t)4095;
    char *d = (char *)dst;
    const char *s = (const char *)src;
    size_t own = std::min(per, bytes);
    {
        std::lock_guard<std::mutex> l(mu_);
        for (size_t off = own; off < bytes; off += per) {
            jobs_.push_back(Job{d + off, s + off, std::min(per, bytes - off)});
            pending_++;
        }
    }
    cv_.notify_all();
    copy_nt(d, s, own);
    std::unique_lock<std::mutex> l(mu_);
    done_cv_.wait(l, [this] { return pending_ == 0; });
}

// ---- HostStager -------------------------------------------------------------------------------
int HostStager::init(eut_field *fld, unsigned threads_per_pool)
{
    field = fld;
    const int dev = fld->plan->device;
    device = dev;
    size_t mb = 8;
    if (const char *e = getenv("EUT_STAGE_SLOT_MB")) mb = (size_t)std::min(256, std::max(1, atoi(e)));
    slot_bytes = mb << 20;
    const std::vector<int> cpus = device_local_cpus(dev);
    // the slots are allocated (and thereby placed) by a thread bound to the CPUs next to the GPU
    int rc = EUT_OK;
    std::thread alloc([&] {
        bind_thread(cpus);
        if (cudaSetDevice(dev) != cudaSuccess) { rc = EUT_ERR_CUDA; return; }
        for (int i = 0; i < kSlots && rc == EUT_OK; i++) {
            if (cudaHostAlloc((void **)&up_slot[i], slot_bytes, cudaHostAllocDefault) != cudaSuccess ||
                cudaHostAlloc((void **)&dn_slot[i], slot_bytes, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&up_ev[i], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&dn_ev[i], cudaEventDisableTiming) != cudaSuccess)
                rc = EUT_ERR_ALLOC;
        }
    });
    alloc.join();
    if (rc != EUT_OK) { set_error("pinned staging slots: allocation failed (%s)", cudaGetErrorString(cudaGetLastError())); return rc; }
    const size_t nc = (size_t)std::max<int64_t>(fld->C, 1);
    EUT_CUDA(cudaMalloc((void **)&d_cols, 2 * nc * sizeof(int32_t)));
    EUT_CUDA(cudaHostAlloc((void **)&h_cols, 2 * nc * sizeof(int32_t), cudaHostAllocDefault));
    fill = new CopyPool(threads_per_pool, cpus);
    drain = new CopyPool(threads_per_pool, cpus);
    downloader = std::thread([this, cpus] { bind_thread(cpus); run_downloader(); });
    return EUT_OK;
}

HostStager::~HostStager()
{
    if (downloader.joinable()) {
        { std::lock_guard<std::mutex> l(mu); stop = true; }
        cv.notify_all();
        downloader.join();
    }
    delete fill;
    delete drain;
    cudaSetDevice(device);
    cudaFree(d_cols);
    if (h_cols) cudaFreeHost(h_cols);
    for (int i = 0; i < kSlots; i++) {
        if (up_slot[i]) cudaFreeHost(up_slot[i]);
        if (dn_slot[i]) cudaFreeHost(dn_slot[i]);
        if (up_ev[i]) cudaEventDestroy(up_ev[i]);
        if (dn_ev[i]) cudaEventDestroy(dn_ev[i]);
    }
}

int HostStager::upload(double *d_dst, int64_t dst_ld, const double *h_src, int64_t src_ld, int64_t rows, int64_t cols,
                       cudaStream_t stream)
{
    const bool dense = (dst_ld == rows && src_ld == rows);
    const int64_t spans = dense ? 1 : cols;
    const size_t span_bytes = (size_t)(dense ? rows * cols : rows) * sizeof(double);
    for (int64_t sp = 0; sp < spans; sp++) {
        char *d = (char *)(d_dst + sp * dst_ld);
        const char *s = (const char *)(h_src + sp * src_ld);
        for (size_t off = 0; off < span_bytes; off += slot_bytes) {
            const size_t n = std::min(slot_bytes, span_bytes - off);
            const int slot = (int)(up_count % kSlots);
            const double t0 = now_ms();
            if (up_used[slot]) EUT_CUDA(cudaEventSynchronize(up_ev[slot]));     // the DMA that last read the slot is done
            const double t1 = now_ms();
            fill->copy(up_slot[slot], s + off, n);
            ms_up_wait += t1 - t0; ms_fill += now_ms() - t1;
            EUT_CUDA(cudaMemcpyAsync(d + off, up_slot[slot], n, cudaMemcpyHostToDevice, stream));
            EUT_CUDA(cudaEventRecord(up_ev[slot], stream));
            up_used[slot] = true;
            up_count++;
        }
    }
    return EUT_OK;
}

std::string HostStager::take_stats()
{
    char b[200];
    snprintf(b, sizeof(b), "fill %.2f ms (+%.2f waiting for a slot), drain %.2f ms (+%.2f waiting for the DMA), %u+%u copier threads",
             ms_fill, ms_up_wait, ms_drain, ms_dn_wait, fill ? fill->size() : 0, drain ? drain->size() : 0);
    ms_fill = ms_up_wait = ms_drain = ms_dn_wait = 0;
    return b;
}

uint64_t HostStager::download(DownJob &&job)
{
    std::lock_guard<std::mutex> l(mu);
    job.ticket = ++issued;
    const uint64_t t = job.ticket;
    queue.push_back(std::move(job));
    cv.notify_all();
    return t;
}

int HostStager::wait_ticket(uint64_t ticket)
{
    std::unique_lock<std::mutex> l(mu);
    idle_cv.wait(l, [&] { return finished >= ticket; });
    if (err) { set_error("%s", err_msg.c_str()); return err; }
    return EUT_OK;
}

int HostStager::wait_idle()
{
    std::unique_lock<std::mutex> l(mu);
    idle_cv.wait(l, [&] { return finished >= issued; });
    if (err) { const int e = err; set_error("%s", err_msg.c_str()); err = 0; return e; }
    return EUT_OK;
}

void HostStager::run_downloader()
{
    cudaSetDevice(device);
    for (;;) {
        DownJob j;
        {
            std::unique_lock<std::mutex> l(mu);
            cv.wait(l, [this] { return stop || !queue.empty(); });
            if (queue.empty()) return;
            j = std::move(queue.front());
            queue.pop_front();
        }
        auto fail = [&](cudaError_t e, const char *what) {
            std::lock_guard<std::mutex> l(mu);
            if (!err) { err = EUT_ERR_CUDA; err_msg = std::string("staged download: ") + what + ": " + cudaGetErrorString(e); }
        };
        eut_field *f = field;
        cudaStream_t stream = f->down_stream;
        const int64_t N = f->plan->N, C = std::max<int64_t>(f->C, 1);
        cudaError_t e = cudaStreamWaitEvent(stream, j.after, 0);          // the substeps of these columns are done
        if (e != cudaSuccess) fail(e, "cudaStreamWaitEvent");
        cudaEventDestroy(j.after);
        struct Flight { int slot; char *dst; size_t n; };
        std::deque<Flight> flight;
        uint64_t piece = 0;
        auto retire = [&]() {
            const Flight fl = flight.front();
            flight.pop_front();
            const double t0 = now_ms();
            const cudaError_t e2 = cudaEventSynchronize(dn_ev[fl.slot]);
            if (e2 != cudaSuccess) { fail(e2, "cudaEventSynchronize"); return; }
            const double t1 = now_ms();
            drain->copy(fl.dst, dn_slot[fl.slot], fl.n);
            ms_dn_wait += t1 - t0; ms_drain += now_ms() - t1;
        };
        const int64_t n_cols = (int64_t)j.cols.size();
        for (int64_t j0 = 0; j0 < n_cols; j0 += f->stage_cols) {
            // un-permute up to stage_cols columns into one half of the device staging area; the kernel and
            // the copies out of the area are all on the download stream, so a half is never rewritten
            // before the copies that read it have completed
            const int64_t nc = std::min(f->stage_cols, n_cols - j0);
            const int half = (int)(dn_chunks++ & 1);
            double *stage = f->d_stage2 + (size_t)half * f->stage_cols * N;
            int32_t *hc = h_cols + (size_t)half * C, *dc = d_cols + (size_t)half * C;
            for (int64_t k = 0; k < nc; k++) hc[k] = j.cols[j0 + k];
            e = cudaMemcpyAsync(dc, hc, nc * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
            if (e != cudaSuccess) { fail(e, "cudaMemcpyAsync"); break; }
            if (launch_permute_out(f, stage, N, dc, f->d_cur, nc, stream) != EUT_OK) { fail(cudaGetLastError(), "un-permute launch"); break; }
            const bool dense = (j.dst_ld == N);
            const int64_t spans = dense ? 1 : nc;
            const size_t span_bytes = (size_t)(dense ? N * nc : N) * sizeof(double);
            for (int64_t sp = 0; sp < spans; sp++) {
                const char *s = (const char *)(stage + sp * N);
                char *d = (char *)(j.h_dst + (j0 + sp) * j.dst_ld);
                for (size_t off = 0; off < span_bytes; off += slot_bytes) {
                    const size_t n = std::min(slot_bytes, span_bytes - off);
                    if ((int)flight.size() >= kSlots - 1) retire();        // keep a few DMAs queued while one slot drains
                    const int slot = (int)(piece++ % kSlots);
                    e = cudaMemcpyAsync(dn_slot[slot], s + off, n, cudaMemcpyDeviceToHost, stream);
                    if (e == cudaSuccess) e = cudaEventRecord(dn_ev[slot], stream);
                    if (e != cudaSuccess) { fail(e, "cudaMemcpyAsync"); continue; }
                    flight.push_back(Flight{slot, d + off, n});
                }
            }
        }
        while (!flight.empty()) retire();
        {
            std::lock_guard<std::mutex> l(mu);
            finished = j.ticket;
        }
        idle_cv.notify_all();
    }
}

}  // namespace eut
