// One-shot tier: the two Rcpp entry points call for call (src/diffChange.cpp:6-35, :38-69), host
// buffers in, host buffer out, on one GPU or -- eut_diffchange_par_multi -- on several from ONE call:
// the reference parallelises the compounds of one call over OpenMP threads (src/diffChange.cpp:44-45);
// here the active columns are dealt out to the devices in contiguous blocks, one host thread per
// device drives that device's column pipeline (H2D | substeps | D2H), the graph is replicated once
// per device and the results land in the caller's single conc_out.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <functional>
#include <future>
#include <limits>
#include <memory>
#include <mutex>
#include <thread>

#include "common.h"
#include "hoststage.h"

using namespace eut;

namespace {

// R hands mat.in / mat.out over with every call (R/growthEnvironment.R:385-390).  The graph analysis
// and its device tables are kept and recognised EXACTLY: the arrays of a cached plan are kept on the
// host and compared byte for byte (a parallel memcmp of 83 MB at 1 M points costs about what a hash of
// it did, and cannot collide).
struct GraphKey {
    std::vector<int32_t> adj, nn;
    int64_t n_edges = 0, n_nn = 0;
};

bool same_bytes(const void *a, const void *b, size_t n)
{
    const size_t kSlice = 4u << 20;
    if (n <= 2 * kSlice) return std::memcmp(a, b, n) == 0;
    const size_t n_slices = (n + kSlice - 1) / kSlice;
    const unsigned n_thr = std::min<unsigned>(16, std::max(1u, host_thread_budget()));
    std::atomic<bool> differ{false};
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (size_t s = next++; s < n_slices && !differ.load(std::memory_order_relaxed); s = next++) {
            const size_t off = s * kSlice;
            if (std::memcmp((const char *)a + off, (const char *)b + off, std::min(kSlice, n - off)) != 0) differ = true;
        }
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n_thr; t++) pool.emplace_back(work);
    work();
    for (auto &th : pool) th.join();
    return !differ;
}

bool same_graph(const GraphKey &k, const int32_t *adj, int64_t n_edges, const int32_t *nn, int64_t n_nn)
{
    if (k.n_edges != n_edges || k.n_nn != n_nn) return false;
    return same_bytes(k.nn.data(), nn, (size_t)n_nn * 2 * sizeof(int32_t)) &&
           same_bytes(k.adj.data(), adj, (size_t)n_edges * 2 * sizeof(int32_t));
}

struct CachedPlan {
    std::shared_ptr<GraphKey> key;
    int64_t n_rows = 0;
    int device = 0;
    eut_plan *plan = nullptr;
    eut_field *field = nullptr, *field1 = nullptr;
    int64_t field_cols = 0;
    uint64_t stamp = 0;
    const void *last_adj = nullptr, *last_nn = nullptr;   // the arrays of the last call (a hint, never trusted)
};

std::mutex g_mu;                                   // the cache list and its stamps
std::mutex g_dev_mu[64];                           // one call at a time per device
std::vector<std::unique_ptr<CachedPlan>> g_cache;
uint64_t g_stamp = 0;
constexpr size_t kMaxCachedPerDevice = 4;

void drop(CachedPlan &c)
{
    if (c.field) field_destroy(c.field);
    if (c.field1) field_destroy(c.field1);
    if (c.plan) { plan_free(c.plan); delete c.plan; }
    c.field = c.field1 = nullptr; c.plan = nullptr;
}

CachedPlan *find_cached(const std::shared_ptr<GraphKey> &key, int64_t n_rows, int device)
{
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &c : g_cache)
        if (c->key == key && c->n_rows == n_rows && c->device == device) return c.get();
    return nullptr;
}

// the plan of (key, device), built on first use (caller holds the device's mutex)
int get_plan(const std::shared_ptr<GraphKey> &key, int64_t n_rows, int device, CachedPlan **out)
{
    if (CachedPlan *c = find_cached(key, n_rows, device)) { *out = c; return EUT_OK; }
    std::unique_ptr<CachedPlan> c(new CachedPlan());
    c->key = key; c->n_rows = n_rows; c->device = device;
    c->plan = new eut_plan();
    const int rc = plan_build(c->plan, key->adj.data(), key->adj.data() + key->n_edges, key->n_edges, key->nn.data(),
                              key->nn.data() + key->n_nn, key->n_nn, n_rows, device, EUT_ORDER_AUTO);
    if (rc != EUT_OK) { plan_free(c->plan); delete c->plan; return rc; }
    std::lock_guard<std::mutex> lock(g_mu);
    size_t mine = 0;
    for (auto &e : g_cache) mine += (e->device == device);
    if (mine >= kMaxCachedPerDevice) {
        auto it = g_cache.end();
        for (auto j = g_cache.begin(); j != g_cache.end(); ++j)
            if ((*j)->device == device && (it == g_cache.end() || (*j)->stamp < (*it)->stamp)) it = j;
        drop(**it);
        g_cache.erase(it);
    }
    c->stamp = ++g_stamp;
    *out = c.get();
    g_cache.push_back(std::move(c));
    return EUT_OK;
}

constexpr int kRetry = -1000;   // internal: the speculated plan was not the right one

// The column pipeline on the plan / field of `cp`.  `verify` (may be empty) is called once, after the
// first chunks have been queued and before anything is written to conc_out; when it returns false
// the streams are drained and kRetry is returned (nothing has been written to conc_out).
int run_pipeline(CachedPlan *cp, const double *conc_in, int64_t n_rows, int64_t n_cols, const int32_t *niter,
                 const int32_t *indvar, const std::vector<int64_t> &act, double *conc_out,
                 const std::function<bool()> &verify, unsigned copy_threads, eut_field **used)
{
    // one field per plan, re-created when the column count changes; single-column calls (the carry-over
    // redo of eut_diffchange) get their own small field instead of flipping the big one back and forth
    eut_field *f = nullptr;
    if (n_cols == 1 && cp->field && cp->field_cols != 1) {
        if (!cp->field1) EUT_TRY(field_create(&cp->field1, cp->plan, 1));
        f = cp->field1;
    } else {
        if (!cp->field || cp->field_cols != n_cols) {
            if (cp->field) field_destroy(cp->field);
            cp->field = nullptr;
            EUT_TRY(field_create(&cp->field, cp->plan, n_cols));
            cp->field_cols = n_cols;
        }
        f = cp->field;
    }
    f->copy_threads = copy_threads;
    *used = f;
    // EUT_ONESHOT_STREAM=1: column streaming (field.cu: one launch at a time over all columns that have
    // arrived).  Measured equal to the chunked pipeline below (157 vs 160 G updates/s): what both pay
    // for is the fixed cost of a launch over few columns (~26 us), and the chunked one replays CUDA graphs.
    static const bool streaming = getenv("EUT_ONESHOT_STREAM") != nullptr;
    if (streaming && act.size() >= 2) {
        std::vector<int32_t> scols(act.size()), snit(act.size());
        for (size_t a = 0; a < act.size(); a++) { scols[a] = indvar[act[a]]; snit[a] = niter[act[a]]; }
        const int rc = field_stream_columns(f, conc_in, conc_out, n_rows, scols.data(), snit.data(), (int64_t)act.size(), verify);
        return rc == -1000 ? kRetry : rc;
    }
    // Pipeline over chunks of active columns: H2D(g+1) | substeps(g) | D2H(g-1) on copy, compute and
    // download streams.  The call lasts  H2D(first chunk) + all substeps + D2H(last chunk), and a launch
    // over few columns is less efficient than one over many (tools/cols_scaling.py).
    // Two compute streams ("lanes"), chunks alternate between them: a launch over few columns has
    // few CTAs (tiles x column slices) and ends in a partial wave; with a second chunk in flight the
    // other lane's CTAs fill the SMs that partial wave leaves idle.  Every chunk gets its own range
    // of the field's launch tables (act_base).  EUT_ONESHOT_LANES=1 restores the single stream.
    static const int n_lanes = [] { const char *e = getenv("EUT_ONESHOT_LANES"); return (e && atoi(e) == 1) ? 1 : 2; }();
    if (n_lanes == 2 && !f->lane[0]) EUT_CUDA(cudaStreamCreateWithFlags(&f->lane[0], cudaStreamNonBlocking));
    cudaStream_t lanes[2] = {f->stream, n_lanes == 2 ? f->lane[0] : f->stream};
    struct Restore {
        eut_field *f; cudaStream_t s;
        ~Restore() { f->stream = s; f->act_base = 0; f->no_split = false; }
    } restore{f, f->stream};
    f->no_split = true;
    const int64_t chunk = f->stage_cols;
    std::vector<size_t> cuts;
    // EUT_ONESHOT_CHUNKS=4,12,32,12,4: explicit chunk sizes (used when they add up to the active columns)
    if (const char *e = getenv("EUT_ONESHOT_CHUNKS")) {
        std::vector<size_t> sizes;
        size_t tot = 0;
        for (const char *q = e; *q;) {
            char *end = nullptr;
            const long v = strtol(q, &end, 10);
            if (end == q) break;
            if (v > 0 && v <= chunk) { sizes.push_back((size_t)v); tot += (size_t)v; }
            q = (*end == ',') ? end + 1 : end;
        }
        if (tot == act.size()) {
            cuts.push_back(0);
            for (size_t sz : sizes) cuts.push_back(cuts.back() + sz);
        }
    }
    if (cuts.empty()) {
        // Default: ramp up to a peak chunk of a quarter of the columns and down again, in steps of a
        // quarter of the peak (64 columns: 4, 8, 12, 16, 12, 8, 4) -- short ends for a short H2D head
        // and D2H tail, and both lanes get the same share.  Measured on 1 M points x 64 columns x 60
        // substeps: 19.9 ms per call against 23.8 ms on one stream with chunks 4, 12, 32, 12, 4.
        // Anything below a quarter of the staging slot (64 MB) goes as one chunk.
        const size_t n = act.size();
        std::vector<size_t> sizes;
        size_t peak = (size_t)chunk;
        auto middle = [&](size_t m) {
            if (m == 0) return;
            const size_t parts = (m + peak - 1) / peak;
            for (size_t k = 0; k < parts; k++) sizes.push_back(m / parts + (k < m % parts ? 1 : 0));
        };
        if (n * 4 > (size_t)chunk) {
            peak = std::min<size_t>((size_t)chunk, std::max<size_t>(4, (n + 3) / 4));
            const size_t r1 = std::max<size_t>(1, peak / 4), r2 = std::max<size_t>(1, peak / 2), r3 = std::max<size_t>(1, 3 * peak / 4);
            if (n > 2 * (r1 + r2 + r3)) {
                sizes.push_back(r1); sizes.push_back(r2); sizes.push_back(r3);
                middle(n - 2 * (r1 + r2 + r3));
                sizes.push_back(r3); sizes.push_back(r2); sizes.push_back(r1);
            } else if (n > 2 * r1) { sizes.push_back(r1); middle(n - 2 * r1); sizes.push_back(r1); }
            else middle(n);
        } else {
            middle(n);
        }
        cuts.push_back(0);
        for (size_t sz : sizes) cuts.push_back(cuts.back() + sz);
    }
    struct Pending { std::vector<int32_t> cols; bool contiguous; cudaStream_t lane; };
    std::vector<Pending> pending;
    bool verified = !verify;
    auto download = [&](const Pending &pd) -> int {
        struct Lane {                                   // the un-permute runs on the lane that diffused the chunk
            eut_field *f; cudaStream_t keep;
            ~Lane() { f->stream = keep; }
        } lane{f, f->stream};
        f->stream = pd.lane;
        if (pd.contiguous) {
            EUT_TRY(field_download_async(f, conc_out + (size_t)(pd.cols[0] - 1) * n_rows, n_rows, pd.cols.data(), (int64_t)pd.cols.size()));
        } else {
            for (size_t j = 0; j < pd.cols.size(); j++)
                EUT_TRY(field_download_async(f, conc_out + (size_t)(pd.cols[j] - 1) * n_rows, n_rows, &pd.cols[j], 1));
        }
        return EUT_OK;
    };
    auto check_now = [&]() -> int {
        if (verified) return EUT_OK;
        if (!verify()) {
            cudaStreamSynchronize(f->copy_stream); cudaStreamSynchronize(lanes[0]); cudaStreamSynchronize(lanes[1]);
            field_sync_down(f);
            return kRetry;
        }
        verified = true;
        for (const Pending &pd : pending) EUT_TRY(download(pd));
        pending.clear();
        return EUT_OK;
    };
    std::vector<int32_t> nit;
    for (size_t ci = 0; ci + 1 < cuts.size(); ci++) {
        const size_t a0 = cuts[ci], a1 = cuts[ci + 1];
        if (a1 <= a0) continue;
        Pending pd;
        pd.lane = lanes[ci & 1];
        f->stream = pd.lane;
        f->act_base = (int64_t)a0;
        nit.clear();
        for (size_t a = a0; a < a1; a++) { pd.cols.push_back(indvar[act[a]]); nit.push_back(niter[act[a]]); }
        // host column j of this chunk is conc_in column cols[j]-1: upload one column at a time when
        // they are not contiguous, as one 2-D copy when they are
        pd.contiguous = true;
        for (size_t j = 1; j < pd.cols.size(); j++) pd.contiguous &= (pd.cols[j] == pd.cols[j - 1] + 1);
        if (pd.contiguous) {
            EUT_TRY(field_upload(f, conc_in + (size_t)(pd.cols[0] - 1) * n_rows, n_rows, pd.cols.data(), (int64_t)pd.cols.size()));
        } else {
            for (size_t j = 0; j < pd.cols.size(); j++)
                EUT_TRY(field_upload(f, conc_in + (size_t)(pd.cols[j] - 1) * n_rows, n_rows, &pd.cols[j], 1));
        }
        EUT_TRY(field_diffuse(f, nit.data(), pd.cols.data(), (int64_t)pd.cols.size()));
        if (!verified && ci >= 2) EUT_TRY(check_now());      // the hash has had three chunks' worth of time; the queue stays fed
        if (verified) EUT_TRY(download(pd));
        else pending.push_back(std::move(pd));
    }
    EUT_TRY(check_now());
    return EUT_OK;
}


// Deal the active columns (positions `act` into niter, ascending column number) out to at most
// n_devices devices: contiguous blocks with about equal substep counts, never an empty block.
// owner[k] = block of act[k]; returns the number of blocks used.
int deal_columns(const int32_t *niter, const std::vector<int64_t> &act, int n_devices, int *owner)
{
    int64_t total = 0;
    for (int64_t a : act) total += niter[a];
    const int n_used = (int)std::min<int64_t>(n_devices, (int64_t)act.size());
    size_t pos = 0;
    int64_t done = 0;
    for (int d = 0; d < n_used; d++) {
        const int64_t goal = (total * (d + 1) + n_used - 1) / n_used;
        const size_t first = pos;
        // a block takes columns while it stays within its share of the work, always at least one, and
        // always leaves one column for every block after it; the last block takes the rest
        while (pos < act.size() && act.size() - pos > (size_t)(n_used - 1 - d) &&
               (pos == first || d == n_used - 1 || done + niter[act[pos]] <= goal)) {
            done += niter[act[pos]];
            owner[pos++] = d;
        }
    }
    while (pos < act.size()) owner[pos++] = n_used - 1;
    return n_used;
}

// What one device does for one call: its block of the active columns through its pipeline.
struct DeviceJob {
    int device = 0;
    std::vector<int64_t> act;        // positions in niter / indvar, ascending column
    int rc = EUT_OK;
    std::string err;
};

// Shared body: `niter[i]` substeps on 1-based column `indvar[i]`; all other columns pass through.
int oneshot_run(const int32_t *adj, int64_t n_edges, const int32_t *nn, int64_t n_nn,
                const double *conc_in, int64_t n_rows, int64_t n_cols, const int32_t *niter,
                const int32_t *indvar, int64_t n_indvar, double *conc_out, const int *devices, int n_devices)
{
    if (n_rows < 0 || n_cols < 0 || n_edges < 0 || n_nn < 0 || n_indvar < 0 ||
        (n_rows * n_cols > 0 && (!conc_in || !conc_out)) || (n_edges > 0 && !adj) ||
        (n_nn > 0 && !nn) || (n_indvar > 0 && (!niter || !indvar))) {
        set_error("one-shot call: NULL pointer or negative size");
        return EUT_ERR_ARG;
    }
    if (conc_in == conc_out && n_rows * n_cols > 0) {
        set_error("one-shot call: conc_out must not alias conc_in (R value semantics)");
        return EUT_ERR_ARG;
    }
    if (n_devices < 1 || !devices || n_devices > 64) { set_error("one-shot call: empty device list"); return EUT_ERR_ARG; }
    for (int d = 0; d < n_devices; d++) {
        EUT_TRY(ensure_device(devices[d]));
        for (int e = 0; e < d; e++)
            if (devices[e] == devices[d]) { set_error("one-shot call: device %d listed twice", devices[d]); return EUT_ERR_ARG; }
    }
    const char *trace_env = getenv("EUT_TRACE");
    const bool trace = trace_env && *trace_env;
    const double t_begin = now_ms();
    // active columns (validated before any work is queued), ascending column number
    std::vector<int64_t> act;
    std::vector<uint8_t> is_act(std::max<int64_t>(n_cols, 1), 0);
    for (int64_t i = 0; i < n_indvar; i++) {
        if (niter[i] <= 0) continue;
        if (indvar[i] < 1 || indvar[i] > n_cols) {
            set_error("index out of bounds: column %d of a %lld-column matrix", indvar[i], (long long)n_cols);
            return EUT_ERR_INDEX;
        }
        if (is_act[indvar[i] - 1]) {
            set_error("column %d listed twice in indvar", indvar[i]);
            return EUT_ERR_ARG;
        }
        is_act[indvar[i] - 1] = 1;
        act.push_back(i);
    }
    // the graph is validated even when nothing diffuses?  No: the reference never touches adjmat then
    // (src/diffChange.cpp:48-55 sits inside the substep loop), so neither do we
    std::vector<DeviceJob> jobs;
    if (!act.empty()) {
        if (n_devices > 1) std::stable_sort(act.begin(), act.end(), [&](int64_t a, int64_t b) { return indvar[a] < indvar[b]; });
        std::vector<int> owner(act.size());
        const int n_used = deal_columns(niter, act, n_devices, owner.data());
        jobs.resize(n_used);
        for (int d = 0; d < n_used; d++) jobs[d].device = devices[d];
        for (size_t k = 0; k < act.size(); k++) jobs[owner[k]].act.push_back(act[k]);
    }
    const unsigned copy_threads = std::min(8u, std::max(1u, host_thread_budget() / (2u * (unsigned)std::max<size_t>(1, jobs.size()))));

    // Which graph?  When the caller hands over the same arrays as last time -- R does, every round --
    // the byte comparison runs on other threads while the first chunks are already uploaded and diffused
    // on the plans used last time; conc_out is not touched before the comparison has confirmed the guess.
    std::shared_ptr<GraphKey> guess;
    if (!jobs.empty()) {
        std::lock_guard<std::mutex> lock(g_mu);
        uint64_t best = 0;
        for (auto &c : g_cache)
            if (c->last_adj == adj && c->last_nn == nn && c->key->n_edges == n_edges && c->key->n_nn == n_nn &&
                c->n_rows == n_rows && c->stamp >= best) { best = c->stamp; guess = c->key; }
    }
    const bool speculate = guess && !getenv("EUT_NO_SPECULATION");
    double compare_ms = 0;
    std::shared_future<std::shared_ptr<GraphKey>> final_key;
    if (!jobs.empty()) {
        final_key = std::async(std::launch::async, [&, guess]() -> std::shared_ptr<GraphKey> {
            struct Stamp { double &ms, t0; ~Stamp() { ms = now_ms() - t0; } } stamp{compare_ms, now_ms()};
            if (guess && same_graph(*guess, adj, n_edges, nn, n_nn)) return guess;
            std::vector<std::shared_ptr<GraphKey>> seen;
            {
                std::lock_guard<std::mutex> lock(g_mu);
                for (auto &c : g_cache)
                    if (c->key != guess && std::find(seen.begin(), seen.end(), c->key) == seen.end()) seen.push_back(c->key);
            }
            for (auto &k : seen)
                if (same_graph(*k, adj, n_edges, nn, n_nn)) return k;
            auto k = std::make_shared<GraphKey>();
            k->n_edges = n_edges; k->n_nn = n_nn;
            k->adj.assign(adj, adj + 2 * n_edges);
            k->nn.assign(nn, nn + 2 * n_nn);
            return k;
        }).share();
    }
    auto run_device = [&](DeviceJob &job) {
        std::lock_guard<std::mutex> dev_lock(g_dev_mu[job.device & 63]);
        auto body = [&]() -> int {
            EUT_TRY(ensure_device(job.device));
            // this device's column span [lo, hi] of the caller's matrix becomes columns 1..hi-lo+1 of its field
            int32_t lo = indvar[job.act.front()], hi = lo;
            for (int64_t a : job.act) { lo = std::min(lo, indvar[a]); hi = std::max(hi, indvar[a]); }
            if (jobs.size() == 1) { lo = 1; hi = (int32_t)n_cols; }
            std::vector<int32_t> iv(n_indvar);
            for (int64_t i = 0; i < n_indvar; i++) iv[i] = indvar[i] - (lo - 1);
            const double *in = conc_in + (size_t)(lo - 1) * n_rows;
            double *out = conc_out + (size_t)(lo - 1) * n_rows;
            const int64_t span = hi - lo + 1;
            int rc = kRetry;
            eut_field *f = nullptr;
            CachedPlan *cp = speculate ? find_cached(guess, n_rows, job.device) : nullptr;
            if (cp && job.act.size() > 1)
                rc = run_pipeline(cp, in, n_rows, span, niter, iv.data(), job.act, out,
                                  [&]() { return final_key.get() == guess; }, copy_threads, &f);
            if (rc == kRetry) {
                EUT_TRY(get_plan(final_key.get(), n_rows, job.device, &cp));
                rc = run_pipeline(cp, in, n_rows, span, niter, iv.data(), job.act, out, std::function<bool()>(), copy_threads, &f);
            }
            if (rc != EUT_OK) return rc;
            {
                std::lock_guard<std::mutex> lock(g_mu);
                cp->stamp = ++g_stamp;
                cp->last_adj = adj; cp->last_nn = nn;
            }
            const double t_enq = now_ms();
            EUT_CUDA(cudaStreamSynchronize(f->copy_stream));
            const double t_up = now_ms();
            EUT_CUDA(cudaStreamSynchronize(f->stream));
            if (f->lane[0]) EUT_CUDA(cudaStreamSynchronize(f->lane[0]));
            const double t_comp = now_ms();
            EUT_TRY(field_sync_down(f));
            if (trace)
                fprintf(stderr, "[eut] one-shot, device %d: all queued at %.2f ms, uploads done %.2f, substeps done %.2f, results home %.2f\n",
                        job.device, t_enq - t_begin, t_up - t_begin, t_comp - t_begin, now_ms() - t_begin);
            if (trace && f->stager) fprintf(stderr, "[eut] one-shot, device %d, staged transfers: %s\n", job.device, f->stager->take_stats().c_str());
            return EUT_OK;
        };
        job.rc = body();
        if (job.rc != EUT_OK) job.err = last_error_cstr();
    };
    std::vector<std::thread> workers;
    for (size_t d = 1; d < jobs.size(); d++) workers.emplace_back([&, d] { run_device(jobs[d]); });
    // columns the call does not touch are returned as they came (src/diffChange.cpp:34, :68)
    auto passthrough = [&]() {
        for (int64_t c = 0; c < n_cols; c++)
            if (!is_act[c] && n_rows > 0)
                std::memcpy(conc_out + (size_t)c * n_rows, conc_in + (size_t)c * n_rows, (size_t)n_rows * sizeof(double));
    };
    std::thread side;
    if (jobs.size() > 1) side = std::thread(passthrough);
    if (!jobs.empty()) run_device(jobs[0]);
    for (auto &w : workers) w.join();
    if (side.joinable()) side.join(); else passthrough();
    if (final_key.valid()) final_key.wait();                 // error paths: never leave the thread behind
    for (auto &j : jobs)
        if (j.rc != EUT_OK) { set_error("device %d: %s", j.device, j.err.c_str()); return j.rc; }
    if (trace) fprintf(stderr, "[eut] one-shot: %zu device(s), %zu active columns, %.2f ms (graph comparison %.2f ms, %u threads)\n", jobs.size(), act.size(), now_ms() - t_begin, compare_ms, std::min(16u, host_thread_budget()));
    return EUT_OK;
}

}  // namespace

extern "C" {

void eut_oneshot_reset(void)
{
    std::lock_guard<std::mutex> lock(g_mu);
    for (auto &c : g_cache) drop(*c);
    g_cache.clear();
}

int eut_diffchange_par(const int32_t *adjmat, int64_t n_edges, const int32_t *nneighbors,
                       int64_t n_nn, const double *conc_in, int64_t n_rows, int64_t n_cols,
                       const int32_t *niter, const int32_t *indvar, int64_t n_indvar, int ncores,
                       double *conc_out, int device)
{
    (void)ncores;
    clear_error();
    return oneshot_run(adjmat, n_edges, nneighbors, n_nn, conc_in, n_rows, n_cols, niter, indvar,
                       n_indvar, conc_out, &device, 1);
}

int eut_deal_columns(const int32_t *niter, const int32_t *indvar, int64_t n_indvar, int n_devices, int32_t *owner)
{
    clear_error();
    if (n_indvar < 0 || n_devices < 1 || (n_indvar > 0 && (!niter || !indvar || !owner))) { set_error("eut_deal_columns: bad argument"); return EUT_ERR_ARG; }
    std::vector<int64_t> act;
    for (int64_t i = 0; i < n_indvar; i++) { owner[i] = -1; if (niter[i] > 0) act.push_back(i); }
    std::stable_sort(act.begin(), act.end(), [&](int64_t a, int64_t b) { return indvar[a] < indvar[b]; });
    std::vector<int> own(act.size());
    if (!act.empty()) deal_columns(niter, act, n_devices, own.data());
    for (size_t k = 0; k < act.size(); k++) owner[act[k]] = own[k];
    return EUT_OK;
}

int eut_diffchange_par_multi(const int32_t *adjmat, int64_t n_edges, const int32_t *nneighbors,
                             int64_t n_nn, const double *conc_in, int64_t n_rows, int64_t n_cols,
                             const int32_t *niter, const int32_t *indvar, int64_t n_indvar, int ncores,
                             double *conc_out, const int *devices, int n_devices)
{
    (void)ncores;
    clear_error();
    std::vector<int> all;
    if (!devices || n_devices <= 0) {                       // NULL / 0: every visible device
        const int n = eut_device_count();
        if (n == 0) return ensure_device(0);
        for (int d = 0; d < n; d++) all.push_back(d);
        devices = all.data(); n_devices = n;
    }
    return oneshot_run(adjmat, n_edges, nneighbors, n_nn, conc_in, n_rows, n_cols, niter, indvar,
                       n_indvar, conc_out, devices, n_devices);
}

int eut_diffchange(const double *adjmat, int64_t n_edges, const double *nneighbors, int64_t n_nn,
                   const double *conc_in, int64_t n_rows, int64_t n_cols, const double *niter,
                   int64_t n_niter, double *conc_out, int device)
{
    clear_error();
    if (n_edges < 0 || n_nn < 0 || n_niter < 0 || (n_edges > 0 && !adjmat) ||
        (n_nn > 0 && !nneighbors) || (n_niter > 0 && !niter)) {
        set_error("eut_diffchange: NULL pointer or negative size");
        return EUT_ERR_ARG;
    }
    // (int) casts of src/diffChange.cpp:20,26 -- out-of-range values are where the reference's
    // bounds check fires; they are mapped to 0 here so that plan_build reports EUT_ERR_INDEX.
    auto to_index = [](double v) -> int32_t {
        if (!(v > -2147483648.0 && v < 2147483648.0)) return 0;
        return (int32_t)v;
    };
    std::vector<int32_t> adj((size_t)n_edges * 2), nn((size_t)n_nn * 2);
    for (int64_t j = 0; j < 2 * n_edges; j++) adj[j] = to_index(adjmat[j]);
    for (int64_t j = 0; j < n_nn; j++) nn[j] = to_index(nneighbors[j]);
    // n_neighbors stays a double in the reference ((double) nn - 1, :26): require it to be an
    // integer that int32 holds, which is what mat.out always contains
    for (int64_t j = 0; j < n_nn; j++) {
        double v = nneighbors[n_nn + j];
        if (!(v == std::floor(v)) || !(std::fabs(v) < 2147483648.0)) {
            set_error("eut_diffchange: nneighbors[%lld,2] = %g is not an integer count", (long long)j + 1, v);
            return EUT_ERR_UNSUPPORTED;
        }
        nn[n_nn + j] = (int32_t)v;
    }
    // `for (int k = 0; k < niter(i); k++)` with a double bound (:14)
    std::vector<int32_t> nit, iv;
    for (int64_t i = 0; i < n_niter; i++) {
        double v = niter[i];
        if (!(v > 0)) continue;  // also NaN
        if (!(v < 2147483647.0)) { set_error("eut_diffchange: niter[%lld] is not finite", (long long)i + 1); return EUT_ERR_ARG; }
        if (i >= n_cols) { set_error("index out of bounds: niter has an entry for column %lld of a %lld-column matrix", (long long)i + 1, (long long)n_cols); return EUT_ERR_INDEX; }
        nit.push_back((int32_t)std::ceil(v));
        iv.push_back((int32_t)(i + 1));
    }
    int rc = oneshot_run(adj.data(), n_edges, nn.data(), n_nn, conc_in, n_rows, n_cols, nit.data(),
                         iv.data(), (int64_t)iv.size(), conc_out, &device, 1);
    if (rc != EUT_OK || iv.size() < 2) return rc;
    // `ychg = ychg * 0` (src/diffChange.cpp:16) is not a zero fill and ychg (:10) is shared by all
    // columns: where the last substep of one column left a non-finite change, the next column
    // starts its inflow sum from NaN at that point.  A change is non-finite exactly where the
    // column's result is (c + y with y = +-Inf or NaN; an overflow of the final addition alone is
    // the one case this test cannot see).  Finite data -- every real round -- never gets past the
    // scan; otherwise the columns after the first affected one are redone in sequence: one
    // substep, NaN at the carried points, the remaining substeps.
    // exponent all ones <=> adding 2^52 to the magnitude bits carries into bit 63 (integer adds and
    // ors only: vectorises, no dependent floating-point chain); columns are scanned by a few threads
    auto any_nonfinite = [&](const double *col) {
        uint64_t acc = 0;
        for (int64_t i = 0; i < n_rows; i++) {
            uint64_t b;
            std::memcpy(&b, col + i, sizeof b);
            acc |= (b & 0x7fffffffffffffffull) + 0x0010000000000000ull;
        }
        return (acc >> 63) != 0;
    };
    std::vector<uint8_t> bad(iv.size(), 0);
    {
        const size_t work = iv.size() * (size_t)n_rows;
        const unsigned nt = work < (1u << 22) ? 1u : std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
        std::atomic<size_t> next{0};
        auto scan = [&]() {
            for (size_t k = next++; k < iv.size(); k = next++)
                bad[k] = any_nonfinite(conc_out + (size_t)(iv[k] - 1) * n_rows) ? 1 : 0;
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < nt; t++) pool.emplace_back(scan);
        scan();
        for (auto &th : pool) th.join();
    }
    size_t first = 0;
    while (first < iv.size() && !bad[first]) first++;
    if (first + 1 >= iv.size()) return rc;
    // "change non-finite <=> result non-finite" needs exactly one mat.out row per point: with missing or
    // duplicate rows a point can hold Inf while its ychg stays finite in the reference
    {
        bool regular = (n_nn == n_rows);
        std::vector<uint8_t> seen((size_t)std::max<int64_t>(n_rows, 1), 0);
        for (int64_t j = 0; regular && j < n_nn; j++) {
            const int32_t id = nn[j];
            if (id < 1 || id > n_rows || seen[id - 1]) regular = false;
            else seen[id - 1] = 1;
        }
        if (!regular) {
            set_error("eut_diffchange: non-finite concentrations on a graph without exactly one mat.out row per point "
                      "(the carry-over of src/diffChange.cpp:16 between columns is not reproduced there)");
            return EUT_ERR_UNSUPPORTED;
        }
    }
    std::vector<double> mid((size_t)n_rows);
    const int32_t one = 1;
    for (size_t j = first + 1; j < iv.size(); j++) {
        const double *prev = conc_out + (size_t)(iv[j - 1] - 1) * n_rows;
        double *dst = conc_out + (size_t)(iv[j] - 1) * n_rows;
        rc = oneshot_run(adj.data(), n_edges, nn.data(), n_nn, conc_in + (size_t)(iv[j] - 1) * n_rows, n_rows, 1,
                         &one, &one, 1, mid.data(), &device, 1);
        if (rc != EUT_OK) return rc;
        for (int64_t i = 0; i < n_rows; i++)
            if (!std::isfinite(prev[i])) mid[i] = std::numeric_limits<double>::quiet_NaN();
        const int32_t rest = nit[j] - 1;
        if (rest > 0) {
            rc = oneshot_run(adj.data(), n_edges, nn.data(), n_nn, mid.data(), n_rows, 1, &rest, &one, 1, dst, &device, 1);
            if (rc != EUT_OK) return rc;
        } else {
            std::memcpy(dst, mid.data(), sizeof(double) * (size_t)n_rows);
        }
    }
    return rc;
}


}  // extern "C"
