// Spatial slab sharding of ONE field over several GPUs of the box, driven from ONE process
// (BASELINE.json configs[4]; SURVEY 8e "by spatial slab").  The reference has no counterpart -- it runs
// on one host -- and its `diffChange` signature carries no coordinates, so the slabs are derived from
// the graph: the tile planner's internal order (compact tiles ordered by coarse level band) is cut into
// contiguous ranges, one per slab.  Slab r then works on a LOCAL graph = its own points plus ghost
// copies of the remote in-neighbours:
//   * local ids follow ascending global id, so raster rows stay runs of consecutive ids and the local
//     plan tiles like the global one;
//   * local mat.in = the rows of the global mat.in whose `to` is owned, in the global row order: every
//     owned point sums exactly the same values in exactly the same order as in the unsharded run, so
//     the result is bit-identical to it;
//   * ghosts have no in-edges and n_neighbors = 1 (outflow weight 0): a local substep copies them
//     forward unchanged, the halo exchange overwrites them with the owner's new value.
// Halo exchange: no pack / send / recv / unpack.  After a substep the OWNER of a boundary point stores
// its new value straight into the ghost cell of the neighbour's field through a peer-mapped pointer
// (NVLink P2P inside one process, cudaDeviceEnablePeerAccess): one small kernel per ordered pair of
// neighbouring slabs, ordered against the two substep kernels it sits between by CUDA events.
// (eutropia_b200/slab.py is the same partition with one process per GPU and NCCL send / recv.)
#include <algorithm>
#include <numeric>
#include <thread>

#include "common.h"
#include "hoststage.h"

namespace eut {

// Small blocks on purpose: a substep CTA of the other column group fills an SM's shared memory and all but
// ~5 k of its registers, and the halo kernel has to slip in NEXT to it (64 threads x <= 32 registers do);
// with 256-thread blocks it waited for a substep CTA to retire -- 25 us per substep on a 4-GPU run.
__global__ void __launch_bounds__(64, 8)
k_halo_push(const double *__restrict__ src, const int64_t src_cs, const int64_t src_bs, double *__restrict__ dst,
            const int64_t dst_cs, const int64_t dst_bs, const int32_t *__restrict__ par,
            const int32_t *__restrict__ cols, const int steps_done, const int32_t *__restrict__ spos,
            const int32_t *__restrict__ dpos, const int n)
{
    const int k = blockIdx.x * 64 + threadIdx.x;
    if (k >= n) return;
    const int c = cols[blockIdx.y];
    // buffer that holds the column after `steps_done` substeps of this call (table: buffer at the start);
    // every slab is at the same substep, so it is the same buffer on both sides
    const int64_t b = (par[blockIdx.y] + steps_done) & 1;
    dst[b * dst_bs + (int64_t)c * dst_cs + dpos[k]] = src[b * src_bs + (int64_t)c * src_cs + spos[k]];
}

}  // namespace eut

using namespace eut;

constexpr int kSlabGroups = 4;

struct eut_slab {
    struct Link {                                // values of `n` points travel from part `src` to part `dst`
        int src = 0, dst = 0;
        int64_t n = 0;
        std::vector<int32_t> src_local, dst_local;   // local ids (0-based) on either side, ascending global id
        int32_t *d_spos = nullptr, *d_dpos = nullptr;  // internal positions, resident on the SOURCE device
        cudaEvent_t done[kSlabGroups] = {};            // per column group
    };
    struct Part {
        int device = 0;
        int64_t n_local = 0, n_own = 0;
        std::vector<int32_t> local_global;       // local id -> global id (0-based), ascending
        std::vector<uint8_t> is_own;
        eut_plan *plan = nullptr;
        eut_field *field = nullptr;
        cudaEvent_t stepped[kSlabGroups] = {}, fork = nullptr, join[kSlabGroups] = {};
        cudaStream_t gs[kSlabGroups] = {};       // the stream of every column group (the field's own, its lanes)
        cudaStream_t ps[kSlabGroups] = {};       // high-priority streams of the halo kernels this slab sends
    };
    int64_t N = 0, C = 0;
    int depth = 1;                               // ghost rings per slab = substeps between two halo exchanges
    std::vector<Part> parts;
    std::vector<Link> links;
    int64_t pushes = 0;
    double build_ms = 0;
    // substep loops seen before, captured as ONE CUDA graph over all slabs, streams and devices
    std::map<uint64_t, cudaGraphExec_t> graphs;
    std::map<uint64_t, int64_t> graph_launches;   // kernels one replay stands for
    std::set<uint64_t> graph_seen;
    cudaEvent_t ev_origin = nullptr;
};

extern "C" int64_t eut_slab_launch_count(const eut_slab *s);

namespace {

void slab_free(eut_slab *s)
{
    for (auto &l : s->links) {
        cudaSetDevice(s->parts[l.src].device);
        cudaFree(l.d_spos); cudaFree(l.d_dpos);
        for (cudaEvent_t e : l.done) if (e) cudaEventDestroy(e);
    }
    for (auto &p : s->parts) {
        if (p.field) field_destroy(p.field);
        if (p.plan) { plan_free(p.plan); delete p.plan; }
        cudaSetDevice(p.device);
        for (cudaEvent_t e : p.stepped) if (e) cudaEventDestroy(e);
        for (cudaStream_t st : p.ps) if (st) cudaStreamDestroy(st);
        if (p.fork) cudaEventDestroy(p.fork);
        for (cudaEvent_t e : p.join) if (e) cudaEventDestroy(e);
    }
    for (auto &g : s->graphs) if (g.second) cudaGraphExecDestroy(g.second);
    if (s->ev_origin) cudaEventDestroy(s->ev_origin);
    delete s;
}

int upload_i32(int device, const std::vector<int32_t> &h, int32_t **d)
{
    EUT_TRY(ensure_device(device));
    EUT_CUDA(cudaMalloc((void **)d, std::max<size_t>(h.size(), 1) * sizeof(int32_t)));
    if (!h.empty()) EUT_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    return EUT_OK;
}

// current values of the first n_act columns of group g (table offset `first`) travel along every link
int push_halos(eut_slab *s, int g, int first, int n_act, int steps_done)
{
    if (n_act == 0) return EUT_OK;
    // timing experiments only (results are wrong): 1 = no halo kernels, 2 = no waits between slabs either
    static const int dbg = [] { const char *e = getenv("EUT_SLAB_DEBUG"); return e ? atoi(e) : 0; }();
    if (dbg & 2) return EUT_OK;
    for (auto &l : s->links) {
        eut_slab::Part &a = s->parts[l.src], &b = s->parts[l.dst];
        if (l.n == 0) continue;
        EUT_TRY(ensure_device(a.device));
        // the source's substep has produced the values; the destination's substep has written its ghost cells
        // (a push that lands earlier would be overwritten)
        EUT_CUDA(cudaStreamWaitEvent(a.ps[g], a.stepped[g], 0));
        EUT_CUDA(cudaStreamWaitEvent(a.ps[g], b.stepped[g], 0));
        dim3 grid((unsigned)((l.n + 63) / 64), (unsigned)n_act, 1);
        const eut_field *fa = a.field;
        if (!(dbg & 1)) k_halo_push<<<grid, 64, 0, a.ps[g]>>>(fa->d_buf, a.plan->Np, fa->C * a.plan->Np, b.field->d_buf, b.plan->Np,
                                                b.field->C * b.plan->Np, fa->d_act + fa->act_cap + first, fa->d_act + first,
                                                steps_done, l.d_spos, l.d_dpos, (int)l.n);
        EUT_CUDA(cudaGetLastError());
        EUT_CUDA(cudaEventRecord(l.done[g], a.ps[g]));
        a.field->launches++;
        s->pushes++;
    }
    // nobody reads its ghost cells before they have arrived; and the sender does not run ahead of its own
    // halo kernels (two substeps on it would overwrite the buffer they read)
    for (auto &l : s->links) {
        if (l.n == 0) continue;
        eut_slab::Part &a = s->parts[l.src], &b = s->parts[l.dst];
        EUT_TRY(ensure_device(b.device));
        EUT_CUDA(cudaStreamWaitEvent(b.gs[g], l.done[g], 0));
        EUT_TRY(ensure_device(a.device));
        EUT_CUDA(cudaStreamWaitEvent(a.gs[g], l.done[g], 0));
    }
    return EUT_OK;
}

int mark_stepped(eut_slab *s, int g)
{
    for (auto &p : s->parts) {
        EUT_TRY(ensure_device(p.device));
        EUT_CUDA(cudaEventRecord(p.stepped[g], p.gs[g]));
    }
    return EUT_OK;
}

}  // namespace

extern "C" {

int eut_slab_create(eut_slab **out, const int32_t *from, const int32_t *to, int64_t n_edges, const int32_t *nn_id,
                    const int32_t *nn_cnt, int64_t n_nn, int64_t n_points, int64_t n_cols, const int *devices,
                    int n_devices)
{
    clear_error();
    if (!out) { set_error("eut_slab_create: out is NULL"); return EUT_ERR_ARG; }
    *out = nullptr;
    if (n_devices < 1 || n_devices > 64 || !devices || n_cols < 1 || n_points < 1 || n_edges < 0 || n_nn < 0 ||
        (n_edges > 0 && (!from || !to)) || (n_nn > 0 && (!nn_id || !nn_cnt))) {
        set_error("eut_slab_create: bad argument");
        return EUT_ERR_ARG;
    }
    const double t0 = now_ms();
    const int W = n_devices;
    const int64_t N = n_points;
    for (int d = 0; d < W; d++) EUT_TRY(ensure_device(devices[d]));
    // peer access between the devices of neighbouring slabs (a device may hold several slabs: the tests
    // run three slabs on one GPU)
    for (int a = 0; a < W; a++)
        for (int b = 0; b < W; b++) {
            if (devices[a] == devices[b]) continue;
            int can = 0;
            EUT_CUDA(cudaDeviceCanAccessPeer(&can, devices[a], devices[b]));
            if (!can) { set_error("eut_slab_create: device %d cannot access device %d (no peer-to-peer path)", devices[a], devices[b]); return EUT_ERR_UNSUPPORTED; }
            EUT_CUDA(cudaSetDevice(devices[a]));
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[a], devices[b], cudaGetErrorString(e)); return EUT_ERR_CUDA; }
            (void)cudaGetLastError();
        }
    // global internal order: the tile planner's (host analysis only)
    std::vector<int32_t> owner(N);
    {
        eut_plan g;
        const int rc = plan_build(&g, from, to, n_edges, nn_id, nn_cnt, n_nn, N, devices[0], EUT_ORDER_TILED | EUT_PLAN_HOST_ONLY);
        if (rc != EUT_OK) return rc;
        std::vector<int64_t> bounds(W + 1);
        for (int r = 0; r <= W; r++) bounds[r] = (int64_t)std::llround((double)N * r / W);
        // positions of a tiled plan may exceed N (padding): rank them first
        std::vector<int32_t> by_pos(N);
        std::iota(by_pos.begin(), by_pos.end(), 0);
        std::sort(by_pos.begin(), by_pos.end(), [&](int32_t a, int32_t b) { return g.perm[a] < g.perm[b]; });
        int r = 0;
        for (int64_t k = 0; k < N; k++) {
            while (k >= bounds[r + 1]) r++;
            owner[by_pos[k]] = r;
        }
    }
    std::vector<int32_t> nn_by_id(N, 1);
    for (int64_t j = 0; j < n_nn; j++) {
        if (nn_id[j] < 1 || nn_id[j] > N) { set_error("index out of bounds: mat.out row %lld id = %d", (long long)j + 1, nn_id[j]); return EUT_ERR_INDEX; }
        nn_by_id[nn_id[j] - 1] = nn_cnt[j];
    }
    eut_slab *s = new eut_slab();
    s->N = N; s->C = n_cols;
    // ghost rings per slab = substeps between two exchanges.  Two slabs have nobody to wait for but each other and
    // run best with one ring (2 GPUs x 2 M points x 64: 0.444 ms per substep, 0.449 with four); from three slabs on
    // every substep otherwise ends in a wait for the slower of two neighbours, whose own pace depends on ITS
    // neighbours: four rings, a wait every fourth substep (4 GPUs x 2 M points: 0.511 -> 0.484 ms per substep).
    const int depth = [&] { const char *e = getenv("EUT_SLAB_DEPTH"); return e ? std::min(16, std::max(1, atoi(e))) : (W >= 3 ? 4 : 1); }();
    s->depth = depth;
    s->parts.resize(W);
    std::vector<int> rcs(W, EUT_OK);
    std::vector<std::string> errs(W);
    // per slab: local graph, plan, field -- one host thread each
    std::vector<std::vector<int32_t>> g2l(W);
    auto build_part = [&](int r) -> int {
        eut_slab::Part &P = s->parts[r];
        P.device = devices[r];
        // mark: 1 own, 2 .. depth + 1 ghost rings.  Ring j holds the in-neighbours of ring j - 1 that are not marked
        // yet (ring 0 = own points).  With depth 1 the ghosts are plain copies: no in-edges, outflow weight 0, the
        // halo exchange overwrites them after every substep.  With depth K > 1 the rings 1 .. K-1 are LIVE -- they
        // keep their in-edges and weights and are diffused like own points -- and only ring K is a copy: after an
        // exchange all rings are right; one substep later ring K is stale, ring K-1 was computed from it and is
        // wrong after the second substep, and so on inwards -- the own points stay right for K substeps, so the
        // slabs exchange (and wait for each other) only every K-th substep, for (K-1) x one lattice row per
        // side of redundant work.  Every owned point still sums the same values in the same order: bit-identical.
        std::vector<uint8_t> mark(N, 0);
        for (int64_t i = 0; i < N; i++) mark[i] = owner[i] == r ? 1 : 0;
        for (int ring = 1; ring <= depth; ring++)
            for (int64_t j = 0; j < n_edges; j++)
                if (mark[to[j] - 1] == ring && !mark[from[j] - 1]) mark[from[j] - 1] = (uint8_t)(ring + 1);
        int64_t n_in = 0;
        for (int64_t j = 0; j < n_edges; j++)
            if (mark[to[j] - 1] && mark[to[j] - 1] <= depth) n_in++;          // computed points: own and live rings
        std::vector<int32_t> &map = g2l[r];
        map.assign(N, -1);
        std::vector<uint8_t> lmark;
        for (int64_t i = 0; i < N; i++)
            if (mark[i]) { map[i] = (int32_t)P.local_global.size(); P.local_global.push_back((int32_t)i); P.is_own.push_back(mark[i] == 1); lmark.push_back(mark[i]); }
        P.n_local = (int64_t)P.local_global.size();
        P.n_own = std::count(P.is_own.begin(), P.is_own.end(), (uint8_t)1);
        std::vector<int32_t> lf(n_in), lt(n_in), id(P.n_local), cnt(P.n_local);
        int64_t o = 0;
        for (int64_t j = 0; j < n_edges; j++)
            if (mark[to[j] - 1] && mark[to[j] - 1] <= depth) { lf[o] = map[from[j] - 1] + 1; lt[o] = map[to[j] - 1] + 1; o++; }
        for (int64_t k = 0; k < P.n_local; k++) { id[k] = (int32_t)k + 1; cnt[k] = lmark[k] <= depth ? nn_by_id[P.local_global[k]] : 1; }
        P.plan = new eut_plan();
        if (!getenv("EUT_SLAB_LIVE_GHOSTS")) {               // the outermost ring is overwritten by every exchange: no work for it
            P.plan->dead.resize(P.n_local);
            for (int64_t k = 0; k < P.n_local; k++) P.plan->dead[k] = lmark[k] == depth + 1 ? 1 : 0;
        }
        EUT_TRY(plan_build(P.plan, lf.data(), lt.data(), n_in, id.data(), cnt.data(), P.n_local, P.n_local, P.device, EUT_ORDER_AUTO));
        EUT_TRY(field_create(&P.field, P.plan, n_cols));
        P.field->copy_threads = std::max(1u, std::min(8u, host_thread_budget() / (2u * (unsigned)W)));
        for (int g = 0; g < kSlabGroups; g++) {
            EUT_CUDA(cudaEventCreateWithFlags(&P.stepped[g], cudaEventDisableTiming));
            EUT_CUDA(cudaEventCreateWithFlags(&P.join[g], cudaEventDisableTiming));
        }
        EUT_CUDA(cudaEventCreateWithFlags(&P.fork, cudaEventDisableTiming));
        P.gs[0] = P.field->stream;
        for (int g = 1; g < kSlabGroups; g++) {
            if (!P.field->lane[g - 1]) EUT_CUDA(cudaStreamCreateWithFlags(&P.field->lane[g - 1], cudaStreamNonBlocking));
            P.gs[g] = P.field->lane[g - 1];
        }
        int lo = 0, hi = 0;
        EUT_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int g = 0; g < kSlabGroups; g++) EUT_CUDA(cudaStreamCreateWithPriority(&P.ps[g], cudaStreamNonBlocking, hi));
        return EUT_OK;
    };
    {
        std::vector<std::thread> th;
        for (int r = 0; r < W; r++)
            th.emplace_back([&, r] { rcs[r] = build_part(r); if (rcs[r] != EUT_OK) errs[r] = last_error_cstr(); });
        for (auto &t : th) t.join();
    }
    for (int r = 0; r < W; r++)
        if (rcs[r] != EUT_OK) { set_error("slab %d: %s", r, errs[r].c_str()); const int rc = rcs[r]; slab_free(s); return rc; }
    // links: for every ordered pair (a -> b) the points a owns that b keeps a ghost of, ascending global id
    for (int a = 0; a < W; a++)
        for (int b = 0; b < W; b++) {
            if (a == b) continue;
            eut_slab::Link L;
            L.src = a; L.dst = b;
            const eut_slab::Part &B = s->parts[b];
            for (int64_t k = 0; k < B.n_local; k++)
                if (!B.is_own[k] && owner[B.local_global[k]] == a) {
                    L.dst_local.push_back((int32_t)k);
                    L.src_local.push_back(g2l[a][B.local_global[k]]);
                }
            L.n = (int64_t)L.dst_local.size();
            if (L.n == 0) continue;
            std::vector<int32_t> sp(L.n), dp(L.n);
            for (int64_t k = 0; k < L.n; k++) {
                if (L.src_local[k] < 0) { set_error("eut_slab_create: ghost without an owner copy"); slab_free(s); return EUT_ERR_UNSUPPORTED; }
                sp[k] = s->parts[a].plan->perm[L.src_local[k]];
                dp[k] = B.plan->perm[L.dst_local[k]];
            }
            int rc = upload_i32(s->parts[a].device, sp, &L.d_spos);
            if (rc == EUT_OK) rc = upload_i32(s->parts[a].device, dp, &L.d_dpos);
            for (int g = 0; g < kSlabGroups && rc == EUT_OK; g++)
                if (cudaEventCreateWithFlags(&L.done[g], cudaEventDisableTiming) != cudaSuccess) rc = EUT_ERR_CUDA;
            if (rc != EUT_OK) { slab_free(s); return rc; }
            s->links.push_back(std::move(L));
        }
    if (ensure_device(s->parts[0].device) != EUT_OK || cudaEventCreateWithFlags(&s->ev_origin, cudaEventDisableTiming) != cudaSuccess) { slab_free(s); return EUT_ERR_CUDA; }
    s->build_ms = now_ms() - t0;
    *out = s;
    return EUT_OK;
}

int eut_slab_destroy(eut_slab *s)
{
    if (s) slab_free(s);
    return EUT_OK;
}

int eut_slab_get_info(const eut_slab *s, eut_slab_info *info)
{
    if (!s || !info) { set_error("eut_slab_get_info: NULL argument"); return EUT_ERR_ARG; }
    info->n_points = s->N; info->n_cols = s->C; info->n_slabs = (int32_t)s->parts.size();
    info->n_links = (int32_t)s->links.size();
    info->max_local = 0; info->halo_points = 0; info->max_link = 0; info->segmented = 1;
    for (auto &p : s->parts) { info->max_local = std::max(info->max_local, p.n_local); info->segmented &= p.plan->segmented ? 1 : 0; }
    for (auto &l : s->links) { info->halo_points += l.n; info->max_link = std::max(info->max_link, l.n); }
    info->pushes = s->pushes;
    info->depth = s->depth;
    info->build_ms = s->build_ms;
    return EUT_OK;
}

int eut_slab_part(const eut_slab *s, int slab, int32_t *device, int64_t *n_local, int64_t *n_own, void **stream)
{
    if (!s || slab < 0 || slab >= (int)s->parts.size()) { set_error("eut_slab_part: bad slab index"); return EUT_ERR_ARG; }
    const auto &p = s->parts[slab];
    if (device) *device = p.device;
    if (n_local) *n_local = p.n_local;
    if (n_own) *n_own = p.n_own;
    if (stream) *stream = (void *)p.field->stream;
    return EUT_OK;
}

// host (N x C, column-major, leading dimension ld) -> every slab's rows (own + ghosts)
int eut_slab_upload(eut_slab *s, const double *host, int64_t ld)
{
    clear_error();
    if (!s || !host || ld < s->N) { set_error("eut_slab_upload: NULL pointer or ld < N"); return EUT_ERR_ARG; }
    const int W = (int)s->parts.size();
    std::vector<int> rcs(W, EUT_OK);
    std::vector<std::string> errs(W);
    auto one = [&](int r) -> int {
        eut_slab::Part &P = s->parts[r];
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(s->C, (64ll << 20) / std::max<int64_t>(P.n_local, 1)));
        std::vector<double> tmp((size_t)chunk * P.n_local);
        std::vector<int32_t> cols(chunk);
        for (int64_t c0 = 0; c0 < s->C; c0 += chunk) {
            const int64_t nc = std::min(chunk, s->C - c0);
            for (int64_t c = 0; c < nc; c++) {
                const double *src = host + (size_t)(c0 + c) * ld;
                double *dst = tmp.data() + (size_t)c * P.n_local;
                for (int64_t k = 0; k < P.n_local; k++) dst[k] = src[P.local_global[k]];
                cols[c] = (int32_t)(c0 + c + 1);
            }
            EUT_TRY(field_upload(P.field, tmp.data(), P.n_local, cols.data(), nc));
            EUT_CUDA(cudaStreamSynchronize(P.field->copy_stream));
        }
        EUT_CUDA(cudaStreamSynchronize(P.field->stream));
        return EUT_OK;
    };
    std::vector<std::thread> th;
    for (int r = 0; r < W; r++) th.emplace_back([&, r] { rcs[r] = one(r); if (rcs[r] != EUT_OK) errs[r] = last_error_cstr(); });
    for (auto &t : th) t.join();
    for (int r = 0; r < W; r++)
        if (rcs[r] != EUT_OK) { set_error("slab %d: %s", r, errs[r].c_str()); return rcs[r]; }
    return EUT_OK;
}

// every slab's OWN rows -> host (N x C)
int eut_slab_download(eut_slab *s, double *host, int64_t ld)
{
    clear_error();
    if (!s || !host || ld < s->N) { set_error("eut_slab_download: NULL pointer or ld < N"); return EUT_ERR_ARG; }
    const int W = (int)s->parts.size();
    std::vector<int> rcs(W, EUT_OK);
    std::vector<std::string> errs(W);
    auto one = [&](int r) -> int {
        eut_slab::Part &P = s->parts[r];
        const int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(s->C, (64ll << 20) / std::max<int64_t>(P.n_local, 1)));
        std::vector<double> tmp((size_t)chunk * P.n_local);
        std::vector<int32_t> cols(chunk);
        for (int64_t c0 = 0; c0 < s->C; c0 += chunk) {
            const int64_t nc = std::min(chunk, s->C - c0);
            for (int64_t c = 0; c < nc; c++) cols[c] = (int32_t)(c0 + c + 1);
            EUT_TRY(field_download_async(P.field, tmp.data(), P.n_local, cols.data(), nc));
            EUT_TRY(field_sync_down(P.field));
            for (int64_t c = 0; c < nc; c++) {
                double *dst = host + (size_t)(c0 + c) * ld;
                const double *src = tmp.data() + (size_t)c * P.n_local;
                for (int64_t k = 0; k < P.n_local; k++)
                    if (P.is_own[k]) dst[P.local_global[k]] = src[k];
            }
        }
        return EUT_OK;
    };
    std::vector<std::thread> th;
    for (int r = 0; r < W; r++) th.emplace_back([&, r] { rcs[r] = one(r); if (rcs[r] != EUT_OK) errs[r] = last_error_cstr(); });
    for (auto &t : th) t.join();
    for (int r = 0; r < W; r++)
        if (rcs[r] != EUT_OK) { set_error("slab %d: %s", r, errs[r].c_str()); return rcs[r]; }
    return EUT_OK;
}

// diffChangePar semantics on the sharded field: niter[i] substeps on column indvar[i]; after every substep
// the boundary values of the columns that moved travel to the neighbours' ghost cells
int eut_slab_diffuse(eut_slab *s, const int32_t *niter, const int32_t *indvar, int64_t n_indvar)
{
    clear_error();
    if (!s || n_indvar < 0 || (n_indvar > 0 && (!niter || !indvar))) { set_error("eut_slab_diffuse: bad argument"); return EUT_ERR_ARG; }
    std::vector<uint8_t> seen(s->C, 0);
    std::vector<int64_t> order;
    for (int64_t i = 0; i < n_indvar; i++) {
        if (niter[i] <= 0) continue;                 // never dereferenced by the reference (src/diffChange.cpp:48-54)
        if (indvar[i] < 1 || indvar[i] > s->C) { set_error("index out of bounds: indvar[%lld] = %d, field has %lld columns", (long long)i + 1, indvar[i], (long long)s->C); return EUT_ERR_INDEX; }
        if (seen[indvar[i] - 1]) { set_error("eut_slab_diffuse: column %d listed twice in indvar", indvar[i]); return EUT_ERR_ARG; }
        seen[indvar[i] - 1] = 1;
        order.push_back(i);
    }
    if (order.empty()) return EUT_OK;
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return niter[a] > niter[b]; });
    const int n_tot = (int)order.size();
    // Two column groups with independent substep -> halo -> substep pipelines on two streams per slab: a
    // substep of one group cannot start before its halo has arrived, and meanwhile the other group's
    // substep keeps the SMs busy (it also fills the partial last wave of CTAs, like the lanes of a plain
    // field).  Columns are dealt out round robin, so each group keeps the descending substep counts.
    const char *ge = getenv("EUT_SLAB_GROUPS");
    int G = n_tot >= 32 ? 2 : 1;
    if (ge) G = std::max(1, std::min(std::min(kSlabGroups, atoi(ge)), n_tot / 8 > 0 ? n_tot / 8 : 1));
    int gbeg[kSlabGroups + 1];
    gbeg[0] = 0;
    for (int g = 0; g < G; g++) gbeg[g + 1] = gbeg[g] + (n_tot - g + G - 1) / G;
    std::vector<int64_t> dealt;
    for (int g = 0; g < G; g++)
        for (int a = g; a < n_tot; a += G) dealt.push_back(order[a]);
    // launch tables of every slab (as field_diffuse: column, buffer at the start, buffer at the end), rows at
    // [0, n_tot) in the dealt order -- read by the kernels, so a captured loop does not depend on them
    for (auto &p : s->parts) {
        eut_field *f = p.field;
        EUT_TRY(ensure_device(p.device));
        if (n_tot > f->act_cap) { set_error("eut_slab_diffuse: table range out of bounds"); return EUT_ERR_ARG; }
        for (int a = 0; a < n_tot; a++) {
            const int c = indvar[dealt[a]] - 1;
            f->h_act[a] = c;
            f->h_act[f->act_cap + a] = f->cur[c];
            f->h_act[2 * f->act_cap + a] = (f->cur[c] + niter[dealt[a]]) & 1;
        }
        for (int r = 0; r < 3; r++)
            EUT_CUDA(cudaMemcpyAsync(f->d_act + r * f->act_cap, f->h_act.data() + r * f->act_cap, n_tot * sizeof(int32_t),
                                     cudaMemcpyHostToDevice, p.gs[0]));
        // group 1's stream, and the first slab's stream (every other stream starts behind it, and a captured
        // loop is launched into it): after the tables and whatever this field's stream holds
        EUT_CUDA(cudaEventRecord(p.fork, p.gs[0]));
        for (int g = 1; g < kSlabGroups; g++) EUT_CUDA(cudaStreamWaitEvent(p.gs[g], p.fork, 0));
        if (p.gs[0] != s->parts[0].gs[0]) EUT_CUDA(cudaStreamWaitEvent(s->parts[0].gs[0], p.fork, 0));
    }
    struct Restore {                                 // the fields run on their own stream outside this call
        eut_slab *s;
        ~Restore() { for (auto &p : s->parts) { p.field->stream = p.gs[0]; p.field->act_col = p.field->d_act; p.field->act_par = p.field->d_act + p.field->act_cap; p.field->in_lanes = false; } }
    } restore{s};
    const int max_it = niter[dealt[0]];
    const char *se = getenv("EUT_SLAB_STAGGER");
    const bool stagger = se && atoi(se) == 1;
    cudaStream_t origin = s->parts[0].gs[0];
    auto run_loop = [&]() -> int {
        // every stream starts behind the first slab's stream (in a capture: joins it) ...
        EUT_TRY(ensure_device(s->parts[0].device));
        EUT_CUDA(cudaEventRecord(s->ev_origin, origin));
        for (auto &p : s->parts)
            for (int g = 0; g < kSlabGroups; g++)
                if (p.gs[g] != origin) EUT_CUDA(cudaStreamWaitEvent(p.gs[g], s->ev_origin, 0));
        // ghosts up to date before the first substep (an upload or anything else may have changed owned values)
        for (int g = 0; g < G; g++) {
            EUT_TRY(mark_stepped(s, g));
            EUT_TRY(push_halos(s, g, gbeg[g], gbeg[g + 1] - gbeg[g], 0));
        }
        int n_act[kSlabGroups] = {};
        for (int g = 0; g < G; g++) n_act[g] = gbeg[g + 1] - gbeg[g];
        for (int it = 0; it < max_it; it++) {
            for (int g = 0; g < G; g++) {
                while (n_act[g] > 0 && niter[dealt[gbeg[g] + n_act[g] - 1]] <= it) n_act[g]--;
                if (n_act[g] == 0) continue;
                for (auto &p : s->parts) {
                    eut_field *f = p.field;
                    EUT_TRY(ensure_device(p.device));
                    // staggered groups: this group's substep starts when the other group's has finished, so the
                    // halo exchange of one group runs under the substep of the other
                    if (stagger && G >= 2) EUT_CUDA(cudaStreamWaitEvent(p.gs[g], p.stepped[(g + G - 1) % G], 0));
                    f->stream = p.gs[g]; f->in_lanes = (G >= 2) && !stagger;
                    f->act_col = f->d_act + gbeg[g]; f->act_par = f->d_act + f->act_cap + gbeg[g];
                    EUT_TRY(launch_substep(f, n_act[g], it));
                }
                // the slabs exchange (and wait for each other) every `depth` substeps; after the last substep of
                // the call nothing is pushed: the next call starts with an exchange of its own
                if ((it + 1) % s->depth == 0 && it + 1 < max_it) {
                    EUT_TRY(mark_stepped(s, g));
                    EUT_TRY(push_halos(s, g, gbeg[g], n_act[g], it + 1));
                }
            }
        }
        // ... and ends in it: device mirror of "which buffer holds column c", then the joins
        for (auto &p : s->parts) {
            eut_field *f = p.field;
            EUT_TRY(ensure_device(p.device));
            f->stream = p.gs[0];
            for (int g = 1; g < kSlabGroups; g++) {
                EUT_CUDA(cudaEventRecord(p.join[g], p.gs[g]));
                EUT_CUDA(cudaStreamWaitEvent(p.gs[0], p.join[g], 0));
            }
            EUT_TRY(launch_set_cur(f, f->d_act, f->d_act + 2 * f->act_cap, n_tot));
            if (p.gs[0] != origin) {
                EUT_CUDA(cudaEventRecord(p.stepped[0], p.gs[0]));
                EUT_CUDA(cudaStreamWaitEvent(origin, p.stepped[0], 0));
            }
        }
        return EUT_OK;
    };
    // A loop shape that comes back (R diffuses the same columns every round) is captured the second time it
    // is seen -- one graph over all slabs, both streams of each and all devices -- and replayed from then
    // on: one thread issuing ~60 launches, event records and waits per substep for eight GPUs is slower
    // than the GPUs are.
    static const bool no_graphs = getenv("EUT_NO_GRAPHS") != nullptr;
    uint64_t key = 0xcbf29ce484222325ull;
    auto mix = [&](uint64_t v) { key = (key ^ v) * 0x100000001b3ull; };
    for (int a = 0; a < n_tot; a++) mix((uint64_t)niter[dealt[a]]);
    mix((uint64_t)G); mix((uint64_t)stagger);
    for (auto &p : s->parts) { mix((uint64_t)p.field->opt_cpt); mix((uint64_t)p.field->opt_variant); mix((uint64_t)p.field->opt_shape); mix((uint64_t)p.field->opt_debug); }
    const auto hit = s->graphs.find(key);
    EUT_TRY(ensure_device(s->parts[0].device));
    if (hit != s->graphs.end() && hit->second) {
        EUT_CUDA(cudaGraphLaunch(hit->second, origin));
        s->parts[0].field->launches += s->graph_launches[key];
    } else if (hit == s->graphs.end() && !no_graphs && max_it >= 4 && s->graph_seen.count(key) && s->graphs.size() < 16) {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        const int64_t l0 = eut_slab_launch_count(s);
        EUT_CUDA(cudaStreamBeginCapture(origin, cudaStreamCaptureModeThreadLocal));
        const int rc = run_loop();
        s->graph_launches[key] = eut_slab_launch_count(s) - l0;
        EUT_TRY(ensure_device(s->parts[0].device));
        const cudaError_t ce = cudaStreamEndCapture(origin, &graph);
        bool ok = rc == EUT_OK && ce == cudaSuccess && graph;
        if (ok && cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) ok = false;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
            (void)cudaGetLastError();
            s->graphs[key] = nullptr;                        // do not try again
            EUT_TRY(run_loop());
        } else {
            s->graphs[key] = exec;
            EUT_CUDA(cudaGraphLaunch(exec, origin));
        }
    } else {
        s->graph_seen.insert(key);
        EUT_TRY(run_loop());
    }
    // what the slabs' own streams are given next (a download, the next call's tables) comes after the loop,
    // which ends in the first slab's stream
    EUT_TRY(ensure_device(s->parts[0].device));
    EUT_CUDA(cudaEventRecord(s->ev_origin, origin));
    for (auto &p : s->parts)
        if (p.gs[0] != origin) EUT_CUDA(cudaStreamWaitEvent(p.gs[0], s->ev_origin, 0));
    for (auto &p : s->parts) {
        eut_field *f = p.field;
        for (int a = 0; a < n_tot; a++) {
            const int c = indvar[dealt[a]] - 1;
            f->cur[c] = (uint8_t)((f->cur[c] + niter[dealt[a]]) & 1);
            f->h_cur32[c] = f->cur[c];
        }
        f->version++;
    }
    return EUT_OK;
}

int eut_slab_sync(eut_slab *s)
{
    if (!s) return EUT_ERR_ARG;
    for (auto &p : s->parts) {
        EUT_TRY(ensure_device(p.device));
        EUT_CUDA(cudaStreamSynchronize(p.field->copy_stream));
        EUT_CUDA(cudaStreamSynchronize(p.field->stream));
        EUT_TRY(field_sync_down(p.field));
    }
    return EUT_OK;
}

int64_t eut_slab_launch_count(const eut_slab *s)
{
    int64_t n = 0;
    if (s) for (auto &p : s->parts) n += p.field->launches.load();
    return n;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------
// Cells on a slab-sharded field ("cells ... their field lists may touch the halo").
// Every slab keeps, of every cell, the entries that fall on its OWN points: the slab's cell map is built over
// its local points with the ghosts moved out of every cell's reach, so each (cell, field) entry exists on
// exactly one slab, in the key order of the global list.  What is per field point -- the claim rescale, the
// by-(field, compound) sums of the scatter in cell order, NaN -> 0, the clamp -- then happens on the owner
// exactly as on an unsharded field, bit for bit.  What is per CELL spans the slabs a cell touches and is
// combined here: max distance and sum(max - dist) of the production shares once per cell map; the
// accessible amounts (= colSums of the uptake shares) after every gather, by a kernel on every slab's device
// that adds the partial sums of all slabs in slab order through peer-mapped pointers.  A cell inside one
// slab gets the unsharded bits; a cell across a cut differs by the rounding of one extra addition (1e-16).
// ---------------------------------------------------------------------------------------------------------
namespace eut {
__global__ void __launch_bounds__(256)
k_sum_partials(double *__restrict__ out, const double *const *__restrict__ parts, const int n_parts, const int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = parts[0][i];
    for (int p = 1; p < n_parts; p++) s = __dadd_rn(s, parts[p][i]);
    out[i] = s;
}
}  // namespace eut

struct eut_slab_cells {
    eut_slab *slab = nullptr;
    std::vector<eut_cells *> cells;            // one per slab, on the slab's plan
    std::vector<double *> d_total;             // [M x C] per slab: sums over all slabs
    std::vector<const double **> d_ptrs;       // per slab: device table of every slab's partial sums
    int64_t M = 0, total_cap = 0;
};

extern "C" {

int eut_slab_cells_destroy(eut_slab_cells *sc)
{
    if (!sc) return EUT_OK;
    for (size_t r = 0; r < sc->cells.size(); r++) {
        cudaSetDevice(sc->slab->parts[r].device);
        if (r < sc->d_total.size()) cudaFree(sc->d_total[r]);
        if (r < sc->d_ptrs.size()) cudaFree((void *)sc->d_ptrs[r]);
        if (sc->cells[r]) eut_cells_destroy(sc->cells[r]);
    }
    delete sc;
    return EUT_OK;
}

int eut_slab_cells_create(eut_slab_cells **out, eut_slab *slab)
{
    clear_error();
    if (!out || !slab) { set_error("eut_slab_cells_create: NULL argument"); return EUT_ERR_ARG; }
    *out = nullptr;
    eut_slab_cells *sc = new eut_slab_cells();
    sc->slab = slab;
    const size_t W = slab->parts.size();
    sc->cells.assign(W, nullptr);
    sc->d_total.assign(W, nullptr);
    sc->d_ptrs.assign(W, nullptr);
    for (size_t r = 0; r < W; r++) {
        const int rc = eut_cells_create(&sc->cells[r], slab->parts[r].plan);
        if (rc != EUT_OK) { eut_slab_cells_destroy(sc); return rc; }
    }
    *out = sc;
    return EUT_OK;
}

// `pts`: the GLOBAL field points (n_pts x 3, column-major, leading dimension ld), cells as eut_cells_build_lists
int eut_slab_cells_build_lists(eut_slab_cells *sc, const double *pts, int64_t ld, int64_t n_pts, const double *cx, const double *cy,
                               const double *csize, const double *cscv, int64_t n_cells, int dist_mode)
{
    clear_error();
    if (!sc || !pts || n_pts != sc->slab->N || ld < n_pts || n_cells < 0) { set_error("eut_slab_cells_build_lists: bad argument (points must be the field's %lld)", sc ? (long long)sc->slab->N : 0ll); return EUT_ERR_ARG; }
    eut_slab *s = sc->slab;
    const size_t W = s->parts.size();
    sc->M = n_cells;
    for (size_t r = 0; r < W; r++) {
        const eut_slab::Part &P = s->parts[r];
        // local points in local id order; ghosts are the neighbour's to serve: out of every cell's reach
        std::vector<double> lp((size_t)3 * P.n_local);
        const double far = 1e300;
        for (int64_t k = 0; k < P.n_local; k++) {
            const int64_t g = P.local_global[k];
            const bool own = P.is_own[k] != 0;
            lp[k] = own ? pts[g] : far;
            lp[P.n_local + k] = own ? pts[ld + g] : far;
            lp[2 * P.n_local + k] = own ? pts[2 * ld + g] : far;
        }
        EUT_TRY(eut_cells_build_lists(sc->cells[r], lp.data(), P.n_local, cx, cy, csize, cscv, n_cells, dist_mode));
    }
    if (n_cells == 0) return EUT_OK;
    // field.frac ingredients per cell over ALL its entries (R/run_simulation.R:729-730)
    std::vector<double> dmax((size_t)n_cells, -INFINITY), tmp((size_t)n_cells), fsum((size_t)n_cells, 0.0);
    for (size_t r = 0; r < W; r++) {
        EUT_TRY(cells_get_dmax(sc->cells[r], tmp.data()));
        for (int64_t i = 0; i < n_cells; i++) dmax[i] = std::max(dmax[i], tmp[i]);      // NaN distances: not produced by the cell map
    }
    for (size_t r = 0; r < W; r++) {
        EUT_TRY(cells_partial_fsum(sc->cells[r], dmax.data(), tmp.data()));
        for (int64_t i = 0; i < n_cells; i++) fsum[i] += tmp[i];
    }
    for (size_t r = 0; r < W; r++) EUT_TRY(cells_set_fsum(sc->cells[r], fsum.data()));
    return EUT_OK;
}

int eut_slab_cells_claim_rescale(eut_slab_cells *sc)
{
    clear_error();
    if (!sc) return EUT_ERR_ARG;
    for (auto *c : sc->cells) EUT_TRY(eut_cells_claim_rescale(c));       // per field point: the owner sees every claimant
    return EUT_OK;
}

int64_t eut_slab_cells_num_entries(const eut_slab_cells *sc)
{
    int64_t t = 0;
    if (sc) for (auto *c : sc->cells) t += eut_cells_num_entries(c);
    return t;
}

// entries of one slab with GLOBAL field ids (cell_ptr: n_cells + 1; the other arrays: eut_cells_num_entries of that slab)
int eut_slab_cells_get_lists(const eut_slab_cells *sc, int slab_index, int64_t *n_entries, int64_t *cell_ptr, int32_t *field_id,
                             double *field_dist, double *acc_prop)
{
    clear_error();
    if (!sc || slab_index < 0 || slab_index >= (int)sc->cells.size()) { set_error("eut_slab_cells_get_lists: bad slab index"); return EUT_ERR_ARG; }
    const eut_cells *c = sc->cells[slab_index];
    const int64_t t = eut_cells_num_entries(c);
    if (n_entries) *n_entries = t;
    if (!cell_ptr && !field_id && !field_dist && !acc_prop) return EUT_OK;
    EUT_TRY(eut_cells_get_lists(c, cell_ptr, field_id, field_dist, acc_prop));
    if (field_id) {
        const eut_slab::Part &P = sc->slab->parts[slab_index];
        for (int64_t k = 0; k < t; k++) field_id[k] = P.local_global[field_id[k] - 1] + 1;
    }
    return EUT_OK;
}

// accessible amounts of every compound for every cell (M x C, column-major; fmol_host may be NULL)
int eut_slab_cells_gather(eut_slab_cells *sc, double field_vol, double *fmol_host)
{
    clear_error();
    if (!sc) return EUT_ERR_ARG;
    eut_slab *s = sc->slab;
    const size_t W = s->parts.size();
    const int64_t n = sc->M * s->C;
    if (n == 0) return EUT_OK;
    for (size_t r = 0; r < W; r++) EUT_TRY(eut_cells_gather(sc->cells[r], s->parts[r].field, field_vol, nullptr));
    for (size_t r = 0; r < W; r++) { EUT_TRY(ensure_device(s->parts[r].device)); EUT_CUDA(cudaStreamSynchronize(s->parts[r].field->stream)); }
    std::vector<const double *> h_ptrs(W);
    for (size_t r = 0; r < W; r++) h_ptrs[r] = cells_fmol(sc->cells[r]);
    for (size_t r = 0; r < W; r++) {
        EUT_TRY(ensure_device(s->parts[r].device));
        if (n > sc->total_cap || !sc->d_total[r]) {
            cudaFree(sc->d_total[r]); sc->d_total[r] = nullptr;
            EUT_CUDA(cudaMalloc((void **)&sc->d_total[r], (size_t)(n + n / 8 + 16) * sizeof(double)));
        }
        if (!sc->d_ptrs[r]) EUT_CUDA(cudaMalloc((void **)&sc->d_ptrs[r], W * sizeof(double *)));
        cudaStream_t st = s->parts[r].field->stream;
        EUT_CUDA(cudaMemcpyAsync((void *)sc->d_ptrs[r], h_ptrs.data(), W * sizeof(double *), cudaMemcpyHostToDevice, st));
        k_sum_partials<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(sc->d_total[r], sc->d_ptrs[r], (int)W, n);
        EUT_CUDA(cudaGetLastError());
        s->parts[r].field->launches++;
    }
    sc->total_cap = std::max(sc->total_cap, n + n / 8 + 16);
    // every slab has read every partial sum before any of them is replaced by the total
    for (size_t r = 0; r < W; r++) { EUT_TRY(ensure_device(s->parts[r].device)); EUT_CUDA(cudaStreamSynchronize(s->parts[r].field->stream)); }
    for (size_t r = 0; r < W; r++) {
        EUT_TRY(ensure_device(s->parts[r].device));
        EUT_CUDA(cudaMemcpyAsync(cells_fmol(sc->cells[r]), sc->d_total[r], (size_t)n * sizeof(double), cudaMemcpyDeviceToDevice, s->parts[r].field->stream));
        cells_mark_gathered(sc->cells[r], s->parts[r].field, field_vol);
    }
    if (fmol_host) {
        EUT_TRY(ensure_device(s->parts[0].device));
        EUT_CUDA(cudaMemcpyAsync(fmol_host, sc->d_total[0], (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s->parts[0].field->stream));
        EUT_CUDA(cudaStreamSynchronize(s->parts[0].field->stream));
    }
    return EUT_OK;
}

// apply the cells' exchange fluxes to the sharded field and clamp (eut_cells_scatter on every slab, with the
// colSums over all slabs: a gather that is not current for this state of the field is taken first)
int eut_slab_cells_scatter(eut_slab_cells *sc, double field_vol, const uint8_t *is_constant, const int64_t *ex_ptr,
                           const int32_t *ex_cpd, const double *ex_flx)
{
    clear_error();
    if (!sc || !ex_ptr) { set_error("eut_slab_cells_scatter: bad argument"); return EUT_ERR_ARG; }
    eut_slab *s = sc->slab;
    const size_t W = s->parts.size();
    if (sc->M == 0 || ex_ptr[sc->M] == 0) return EUT_OK;
    bool current = true;
    for (size_t r = 0; r < W; r++) current = current && cells_gather_is_current(sc->cells[r], s->parts[r].field, field_vol);
    if (!current) EUT_TRY(eut_slab_cells_gather(sc, field_vol, nullptr));
    for (size_t r = 0; r < W; r++)
        EUT_TRY(eut_cells_scatter(sc->cells[r], s->parts[r].field, field_vol, is_constant, ex_ptr, ex_cpd, ex_flx));
    return EUT_OK;
}

}  // extern "C"
