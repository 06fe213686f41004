// Plan construction: host-side analysis of mat.in / mat.out and upload of the device-resident
// neighbour tables.  The graph is consumed exactly as given (R/growthEnvironment.R:138-149 builds
// it once per simulation); nothing is regenerated from geometry.
#include <algorithm>
#include <chrono>
#include <numeric>
#include <cstdlib>
#include <thread>

#include "common.h"

namespace eut {

static thread_local std::string g_err;

void set_error(const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
void clear_error() { g_err.clear(); }
const char *last_error_cstr() { return g_err.c_str(); }

double now_ms()
{
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

int ensure_device(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device available (%s); eutropia_b200 has no CPU path",
                  e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return EUT_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= n) {
        set_error("device ordinal %d out of range (have %d)", device, n);
        return EUT_ERR_ARG;
    }
    EUT_CUDA(cudaSetDevice(device));
    return EUT_OK;
}

static uint64_t hash_span(const uint8_t *p, size_t n, uint64_t seed)
{
    // 64-bit multiply-rotate over 8-byte words, four independent lanes so the multiplies pipeline
    uint64_t h[4] = {seed ^ 0x9E3779B97F4A7C15ull, seed ^ 0xC2B2AE3D27D4EB4Full,
                     seed ^ 0x165667B19E3779F9ull, seed ^ 0xBF58476D1CE4E5B9ull};
    const size_t nw = n / 8, n4 = nw / 4 * 4;
    for (size_t i = 0; i < n4; i += 4) {
        uint64_t w[4];
        memcpy(w, p + 8 * i, 32);
        for (int l = 0; l < 4; l++) {
            h[l] ^= w[l] * 0xC2B2AE3D27D4EB4Full;
            h[l] = (h[l] << 31) | (h[l] >> 33);
            h[l] *= 0x9E3779B97F4A7C15ull;
        }
    }
    uint64_t r = h[0] ^ (h[1] * 3) ^ (h[2] * 5) ^ (h[3] * 7) ^ (n * 0x9E3779B97F4A7C15ull);
    for (size_t i = n4; i < nw; i++) {
        uint64_t w;
        memcpy(&w, p + 8 * i, 8);
        r = (r ^ w) * 0x9E3779B97F4A7C15ull;
        r = (r << 29) | (r >> 35);
    }
    uint64_t tail = 0;
    memcpy(&tail, p + 8 * nw, n - 8 * nw);
    r ^= tail * 0x165667B19E3779F9ull;
    r ^= r >> 29;
    r *= 0xBF58476D1CE4E5B9ull;
    r ^= r >> 32;
    return r;
}

uint64_t hash_bytes(const void *data, size_t n, uint64_t seed)
{
    // Only used to recognise "same graph as last call" (R hands mat.in / mat.out over again every
    // round, R/growthEnvironment.R:385-390).  Large arrays are hashed in parallel slices.
    const uint8_t *p = static_cast<const uint8_t *>(data);
    const size_t kSlice = 8u << 20;
    if (n <= 2 * kSlice) return hash_span(p, n, seed);
    const size_t n_slices = (n + kSlice - 1) / kSlice;
    std::vector<uint64_t> part(n_slices);
    // up to 16 threads, but not more than this process's share of the host when several ranks of a
    // launcher run on it (LOCAL_WORLD_SIZE is exported by torchrun)
    unsigned share = std::max(1u, std::thread::hardware_concurrency());
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) share = std::max(2u, share / (unsigned)std::max(1, atoi(e)));
    const unsigned n_thr = std::min<unsigned>(16, share);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < n_thr; t++)
        pool.emplace_back([&, t]() {
            for (size_t s = t; s < n_slices; s += n_thr)
                part[s] = hash_span(p + s * kSlice, std::min(kSlice, n - s * kSlice), seed + s);
        });
    for (auto &th : pool) th.join();
    return hash_span(reinterpret_cast<const uint8_t *>(part.data()), part.size() * 8, seed ^ n);
}

template <typename T>
static int upload(T **dptr, const std::vector<T> &h, int64_t *bytes)
{
    size_t n = h.size() ? h.size() : 1;
    EUT_CUDA(cudaMalloc(reinterpret_cast<void **>(dptr), n * sizeof(T)));
    if (!h.empty())
        EUT_CUDA(cudaMemcpy(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *bytes += (int64_t)(n * sizeof(T));
    return EUT_OK;
}

// ---------------------------------------------------------------------------------------------
// Internal point ordering.  Any permutation is arithmetically neutral: the per-destination
// summation order is the mat.in row order stored in the neighbour tables, not the memory order.
// ---------------------------------------------------------------------------------------------
static void order_identity(eut_plan *p)
{
    p->perm.resize(p->N);
    std::iota(p->perm.begin(), p->perm.end(), 0);
}

// ROWBLOCK: derive "rows" (maximal chains i, i+1, ... joined by edges in both directions -- in the
// reference's numbering these are the raster rows of one layer, R/growthEnvironment.R:86-88) and
// order whole rows by a breadth-first sweep over the row adjacency graph, so that the rows a row
// exchanges with (same layer y+-1, adjacent layers) sit next to it in memory instead of a whole
// layer away.  Points keep their order inside a row, so warps still read runs of consecutive ids.
static void order_rowblock(eut_plan *p)
{
    const int64_t N = p->N;
    std::vector<uint8_t> linked(N, 0);  // linked[i]: edge (i+1 -> i) and (i -> i+1) both exist
    auto has_in = [&](int64_t dst, int32_t src) {
        for (int64_t k = p->row_ptr[dst]; k < p->row_ptr[dst + 1]; k++)
            if (p->col[k] == src) return true;
        return false;
    };
    for (int64_t i = 0; i + 1 < N; i++)
        linked[i] = has_in(i, (int32_t)(i + 1)) && has_in(i + 1, (int32_t)i);
    std::vector<int32_t> row_of(N);
    std::vector<int64_t> row_start;
    for (int64_t i = 0; i < N; i++) {
        if (i == 0 || !linked[i - 1]) row_start.push_back(i);
        row_of[i] = (int32_t)row_start.size() - 1;
    }
    const int64_t R = (int64_t)row_start.size();
    row_start.push_back(N);

    // BFS over rows, visiting the neighbours of a row in ascending row id; restart at the lowest
    // unvisited row for disconnected pieces.
    std::vector<uint8_t> seen(R, 0);
    std::vector<int32_t> row_order;
    row_order.reserve(R);
    std::vector<int32_t> nb;
    for (int64_t seed = 0; seed < R; seed++) {
        if (seen[seed]) continue;
        seen[seed] = 1;
        size_t head = row_order.size();
        row_order.push_back((int32_t)seed);
        while (head < row_order.size()) {
            int32_t r = row_order[head++];
            nb.clear();
            for (int64_t i = row_start[r]; i < row_start[r + 1]; i++)
                for (int64_t k = p->row_ptr[i]; k < p->row_ptr[i + 1]; k++) {
                    int32_t rr = row_of[p->col[k]];
                    if (!seen[rr]) { seen[rr] = 1; nb.push_back(rr); }
                }
            std::sort(nb.begin(), nb.end());
            for (int32_t rr : nb) row_order.push_back(rr);
        }
    }
    p->perm.resize(N);
    int64_t pos = 0;
    for (int32_t r : row_order)
        for (int64_t i = row_start[r]; i < row_start[r + 1]; i++) p->perm[i] = (int32_t)pos++;
    p->n_tiles = (int)R;
}

int plan_build(eut_plan *p, const int32_t *from, const int32_t *to, int64_t n_edges,
               const int32_t *nn_id, const int32_t *nn_cnt, int64_t n_nn, int64_t n_points,
               int device, uint32_t flags)
{
    const double t0 = now_ms();
    if (n_edges < 0 || n_nn < 0 || n_points < 0 || (n_edges > 0 && (!from || !to)) ||
        (n_nn > 0 && (!nn_id || !nn_cnt))) {
        set_error("eut_plan_create: NULL pointer or negative size");
        return EUT_ERR_ARG;
    }
    if (n_points >= (int64_t)INT32_MAX - 2 * kPadPoints) {
        set_error("eut_plan_create: %lld points exceed the int32 index range of the reference",
                  (long long)n_points);
        return EUT_ERR_UNSUPPORTED;
    }
    const bool host_only = (flags & EUT_PLAN_HOST_ONLY) != 0;
    if (!host_only) EUT_TRY(ensure_device(device));
    p->host_only = host_only;
    p->device = device;
    p->N = n_points;
    p->E = n_edges;
    const int64_t N = n_points;

    // -- bounds: where the reference's Armadillo accessors would throw ------------------------
    for (int64_t j = 0; j < n_edges; j++) {
        if (from[j] < 1 || from[j] > N || to[j] < 1 || to[j] > N) {
            set_error("index out of bounds: mat.in row %lld = (%d, %d), field has %lld points",
                      (long long)j + 1, from[j], to[j], (long long)N);
            return EUT_ERR_INDEX;
        }
    }
    for (int64_t j = 0; j < n_nn; j++) {
        if (nn_id[j] < 1 || nn_id[j] > N) {
            set_error("index out of bounds: mat.out row %lld id = %d, field has %lld points",
                      (long long)j + 1, nn_id[j], (long long)N);
            return EUT_ERR_INDEX;
        }
    }

    // -- in-neighbour CSR: stable counting sort by `to`, keeps mat.in row order per destination
    p->row_ptr.assign(N + 1, 0);
    for (int64_t j = 0; j < n_edges; j++) p->row_ptr[to[j]]++;  // to is 1-based -> slot to[j]
    for (int64_t i = 0; i < N; i++) p->row_ptr[i + 1] += p->row_ptr[i];
    p->col.resize(n_edges);
    {
        std::vector<int64_t> fill(p->row_ptr.begin(), p->row_ptr.end() - 1);
        for (int64_t j = 0; j < n_edges; j++) p->col[fill[to[j] - 1]++] = from[j] - 1;
    }
    p->max_in = 0;
    for (int64_t i = 0; i < N; i++)
        p->max_in = std::max<int>(p->max_in, (int)(p->row_ptr[i + 1] - p->row_ptr[i]));

    // -- outflow rows per point, mat.out row order -------------------------------------------
    std::vector<int64_t> out_ptr(N + 1, 0);
    for (int64_t j = 0; j < n_nn; j++) out_ptr[nn_id[j]]++;
    bool regular = (n_nn == N);
    for (int64_t i = 0; i < N; i++) {
        if (out_ptr[i + 1] != 1) regular = false;
        out_ptr[i + 1] += out_ptr[i];
    }
    p->regular_out = regular ? 1 : 0;
    std::vector<double> out_w(n_nn);
    {
        std::vector<int64_t> fill(out_ptr.begin(), out_ptr.end() - 1);
        for (int64_t j = 0; j < n_nn; j++)
            out_w[fill[nn_id[j] - 1]++] = (((double)nn_cnt[j]) - 1) * kBDiv;  // src/diffChange.cpp:60
    }

    // -- ordering -----------------------------------------------------------------------------
    const bool ell_ok = regular && p->max_in <= kEllMax && !(flags & EUT_PLAN_FORCE_GENERIC);
    uint32_t order = flags & EUT_ORDER_MASK;
    if (order == EUT_ORDER_AUTO) {
        const char *e = getenv("EUT_ORDER");
        order = !ell_ok ? EUT_ORDER_ROWBLOCK : (e && atoi(e) == EUT_ORDER_TILED ? EUT_ORDER_TILED : EUT_ORDER_SEGMENTS);
    }
    p->tiled = false;
    p->segmented = false;
    if (order == EUT_ORDER_SEGMENTS && N > 0 && ell_ok) {
        SegParams sp;
        if (const char *e = getenv("EUT_SEG_R")) sp.R = atoi(e);
        sp.tile_cols = 8 * sp.R;
        if (const char *e = getenv("EUT_SEG_COLS")) sp.tile_cols = atoi(e);
        // tile capacity: 448 segments (one CTA of 14 consumer warps per SM) once the field is large enough
        // to give every SM several such tiles; 224 (two CTAs of 7 per SM) below that
        sp.max_tile_segs = N >= 400000 ? 448 : 224;
        if (const char *e = getenv("EUT_SEG_TILE")) sp.max_tile_segs = std::min(448, std::max(8, atoi(e)));
        if (const char *e = getenv("EUT_TILE_BAND")) sp.band_levels = atoi(e);
        if (const char *e = getenv("EUT_SEG_PAD")) sp.pad_rows = atoi(e);
        if (const char *e = getenv("EUT_SEG_MAXPTS")) sp.max_tile_points = std::min(4000, std::max(64, atoi(e)));
        if (build_segs(p, sp) == EUT_OK) p->segmented = true;
        else order = EUT_ORDER_TILED;      // not lattice-like enough / tiles too large: first tiled kernel
    } else if (order == EUT_ORDER_SEGMENTS) order = EUT_ORDER_ROWBLOCK;
    if (order == EUT_ORDER_TILED && N > 0 && ell_ok) {
        TileParams tp;
        if (const char *e = getenv("EUT_TILE_COLS")) tp.tile_cols = atoi(e);
        if (const char *e = getenv("EUT_TILE_POINTS")) tp.max_tile_points = std::min(1536, std::max(64, atoi(e)));
        if (const char *e = getenv("EUT_TILE_BAND")) tp.band_levels = atoi(e);
        if (build_tiles(p, tp) == EUT_OK) p->tiled = true;
        else order = EUT_ORDER_ROWBLOCK;   // graph not lattice-like enough: ELL kernel
    } else if (order == EUT_ORDER_TILED) order = EUT_ORDER_ROWBLOCK;
    if (!p->tiled && !p->segmented) {
        if (order == EUT_ORDER_ROWBLOCK && N > 0) order_rowblock(p);
        else { order = EUT_ORDER_IDENTITY; order_identity(p); }
        p->Np = ((N + kPadPoints - 1) / kPadPoints) * kPadPoints;
        if (p->Np == 0) p->Np = kPadPoints;
        p->iperm.assign(p->Np, -1);
        for (int64_t i = 0; i < N; i++) p->iperm[p->perm[i]] = (int32_t)i;
    }
    p->order = (int)order;
    if (host_only) {   // analysis only (CPU tests of the planner): no device tables
        p->ell_w = ell_ok ? std::max(p->max_in, 1) : 0;
        p->build_ms = now_ms() - t0;
        return EUT_OK;
    }

    // -- device tables ------------------------------------------------------------------------
    p->device_bytes = 0;
    EUT_TRY(upload(&p->d_perm, p->perm, &p->device_bytes));
    EUT_TRY(upload(&p->d_iperm, p->iperm, &p->device_bytes));
    if (ell_ok) {
        p->ell_w = std::max(p->max_in, 1);
        std::vector<int32_t> nbr((size_t)p->ell_w * p->Np);
        std::vector<uint8_t> deg(p->Np, 0);
        std::vector<double> w(p->Np, 0.0);
        for (int64_t q = 0; q < p->Np; q++) {
            int32_t i = p->iperm[q];
            int d = 0;
            if (i >= 0) {
                d = (int)(p->row_ptr[i + 1] - p->row_ptr[i]);
                deg[q] = (uint8_t)d;
                w[q] = out_w[out_ptr[i]];
                for (int k = 0; k < d; k++)
                    nbr[(size_t)k * p->Np + q] = p->perm[p->col[p->row_ptr[i] + k]];
            }
            for (int k = d; k < p->ell_w; k++) nbr[(size_t)k * p->Np + q] = (int32_t)q;
        }
        EUT_TRY(upload(&p->d_nbr, nbr, &p->device_bytes));
        EUT_TRY(upload(&p->d_deg, deg, &p->device_bytes));
        EUT_TRY(upload(&p->d_w, w, &p->device_bytes));
        if (p->tiled) {
            EUT_TRY(upload(&p->d_tile_meta, p->tiles.meta, &p->device_bytes));
            EUT_TRY(upload(&p->d_copies, p->tiles.copies, &p->device_bytes));
            EUT_TRY(upload(&p->d_loc16, p->tiles.loc, &p->device_bytes));
            p->tiles.loc.clear(); p->tiles.loc.shrink_to_fit();
            p->tiles.copies.clear(); p->tiles.copies.shrink_to_fit();
        }
        if (p->segmented) {
            SegPlanHost &S = p->segs;
            EUT_TRY(upload(&p->d_seg_meta, S.meta, &p->device_bytes));
            EUT_TRY(upload(&p->d_copies, S.copies, &p->device_bytes));
            std::vector<uint4> tmp(S.segs.size() / 8);   // 24 x uint16 = 3 x uint4 per segment
            if (!tmp.empty()) memcpy(tmp.data(), S.segs.data(), S.segs.size() * sizeof(uint16_t));
            EUT_TRY(upload(&p->d_segs, tmp, &p->device_bytes));
            tmp.assign(S.gens.size() / 8, make_uint4(0, 0, 0, 0));
            if (!tmp.empty()) memcpy(tmp.data(), S.gens.data(), S.gens.size() * sizeof(uint16_t));
            EUT_TRY(upload(&p->d_gens, tmp, &p->device_bytes));
            S.segs.clear(); S.segs.shrink_to_fit(); S.gens.clear(); S.gens.shrink_to_fit();
            S.copies.clear(); S.copies.shrink_to_fit();
        }
    } else {
        p->ell_w = 0;
        std::vector<int64_t> rp(N + 1, 0), op(N + 1, 0);
        std::vector<int32_t> cl(n_edges);
        std::vector<double> ow(n_nn);
        for (int64_t q = 0; q < N; q++) {
            int32_t i = p->iperm[q];
            rp[q + 1] = rp[q] + (p->row_ptr[i + 1] - p->row_ptr[i]);
            op[q + 1] = op[q] + (out_ptr[i + 1] - out_ptr[i]);
            for (int64_t k = p->row_ptr[i], o = rp[q]; k < p->row_ptr[i + 1]; k++, o++)
                cl[o] = p->perm[p->col[k]];
            for (int64_t k = out_ptr[i], o = op[q]; k < out_ptr[i + 1]; k++, o++) ow[o] = out_w[k];
        }
        EUT_TRY(upload(&p->d_row_ptr, rp, &p->device_bytes));
        EUT_TRY(upload(&p->d_col, cl, &p->device_bytes));
        EUT_TRY(upload(&p->d_out_ptr, op, &p->device_bytes));
        EUT_TRY(upload(&p->d_out_w, ow, &p->device_bytes));
    }
    p->build_ms = now_ms() - t0;
    return EUT_OK;
}

void plan_free(eut_plan *p)
{
    if (!p || p->host_only) return;
    cudaSetDevice(p->device);
    cudaFree(p->d_nbr); cudaFree(p->d_deg); cudaFree(p->d_w); cudaFree(p->d_perm);
    cudaFree(p->d_iperm); cudaFree(p->d_row_ptr); cudaFree(p->d_col); cudaFree(p->d_out_ptr);
    cudaFree(p->d_out_w);
    cudaFree(p->d_tile_meta); cudaFree(p->d_copies); cudaFree(p->d_loc16);
    cudaFree(p->d_seg_meta); cudaFree(p->d_segs); cudaFree(p->d_gens);
}

}  // namespace eut
