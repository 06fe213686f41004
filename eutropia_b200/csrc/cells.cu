// Cell <-> field exchange: cell->field lists (R/run_simulation.R:217-241, :591-624), gather
// (R/run_simulation.R:249 + R/scavenge_compounds.R:17-22) and scatter + clamp
// (R/run_simulation.R:689-740, :309-319), all on the device against the resident field.
//
// Determinism: the reference aggregates the per-cell contributions with a group-by sum in table
// (= cell) order.  Instead of atomics the scatter walks, for every touched field point, the list of
// (cell, entry) pairs that reference it in cell order (a transposed copy of the lists built once
// per cell map), so every (field, compound) sum has a fixed order and the result is reproducible
// bit for bit from run to run.
//
// R's `sum` / `colSums` accumulate in long double (x87 80-bit).  The device has no such type; sums
// that R takes in long double are accumulated here as an unevaluated hi+lo pair (TwoSum), which is
// at least as accurate, and rounded once at the end.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <cmath>
#include <numeric>

#include "common.h"

struct eut_cells {
    eut_plan *plan = nullptr;
    int64_t M = 0;          // cells
    int64_t T = 0;          // entries (sum of list lengths)
    int64_t cap_M = 0, cap_T = 0;
    // per entry, cell-major (table order)
    int64_t *d_cell_ptr = nullptr;  // [M+1]
    int32_t *d_fid = nullptr;       // reference field id, 1-based
    int32_t *d_fq = nullptr;        // internal position
    double *d_dist = nullptr;
    double *d_acc = nullptr;
    int32_t *d_ecell = nullptr;     // cell of entry
    // per cell
    double *d_dmax = nullptr, *d_fsum = nullptr;
    // transposed: entries grouped by field position, cell order inside a group
    int32_t *d_t_entry = nullptr;   // [T] entry index
    int32_t *d_t_key = nullptr;     // [T] sorted field positions
    int32_t *d_seg_start = nullptr; // [n_seg+1] group boundaries in d_t_entry
    int32_t *d_seg_fq = nullptr;    // [n_seg] field position of the group
    int64_t n_seg = 0;
    bool transposed = false;
    // gather order: entries of every cell sorted by internal field position
    int32_t *d_g_idx = nullptr;
    int64_t g_cap = 0;
    int32_t *d_cell_order = nullptr;   // cells sorted by the field position of their first entry
    int64_t order_cap = 0;
    bool g_sorted = false;
    // geometry index for eut_cells_build_lists: field points sorted by (z, x, y)
    int64_t n_pts = 0;
    double *d_gx = nullptr, *d_gy = nullptr;  // sorted coordinates
    int32_t *d_gfield = nullptr;              // 1-based field id in sorted order
    int32_t *d_col_start = nullptr;           // runs of equal (z, x): [n_cols_g+1]
    double *d_col_x = nullptr;                // x of each run
    int32_t *d_layer_col = nullptr;           // [n_layers+1] first run of each z
    double *d_layer_z = nullptr;
    int64_t n_cols_g = 0, n_layers = 0;
    uint64_t pts_hash = 0;
    // scratch
    void *d_tmp = nullptr;
    size_t tmp_bytes = 0;
    double *d_fmol = nullptr;   // [M x C]
    int64_t fmol_cap = 0;
    double *d_tfield = nullptr; // [Np x ldc] point-major copy of the field for the gather
    int64_t tfield_cap = 0;
    double *d_pxy = nullptr;    // [2][Np] x, y of the field points by internal position (chemotaxis moments)
    uint64_t pxy_hash = 0;
    double *d_mom = nullptr;    // [7 x M] result of eut_cells_chemotaxis_moments
    int64_t mom_cap = 0;
    double *d_cxy = nullptr;    // [3 x M] cell x, y, sensing radius
    int64_t cxy_cap = 0;
    double *d_exd = nullptr;    // [M x ldc] dense exchange factors per cell (scatter, second generation)
    int64_t exd_cap = 0;
    unsigned long long *d_exmask = nullptr;   // [M x ldc/64 x 2] bit c: the cell takes up / produces column c
    int64_t exmask_cap = 0;
    int64_t *d_ex_ptr = nullptr;
    int32_t *d_ex_cpd = nullptr;
    double *d_ex_flx = nullptr, *d_ex_cs = nullptr;
    int64_t ex_cap = 0, exptr_cap = 0;
    uint8_t *d_const = nullptr;
    int64_t const_cap = 0;
    int64_t launches = 0;
    // scatter, third generation: the transposed lists packed for a point-per-lane walk
    int32_t *d_t_cell = nullptr;     // [T] cell of every transposed entry
    double2 *d_t_af = nullptr;       // [T] (acc.prop, (dmax - dist) / fsum) of every transposed entry
    int64_t tpack_cap = 0;
    bool t_packed = false;
    int32_t *d_ex_ptr32 = nullptr;   // [M+1]
    double2 *d_exrec = nullptr;      // [X] per exchange: factor | code (column, production bit; < 0: skip)
    int64_t exrec_cap = 0, exptr32_cap = 0;
    // the gather's per-cell sums are the colSums the next scatter divides by (R/run_simulation.R:696 takes
    // them from the same fmolPerField matrix): remembered with the state of the field they were taken from
    const eut_field *fmol_field = nullptr;
    uint64_t fmol_version = 0;
    double fmol_vol = 0;
    // the exchange lists of a scatter call travel on the field's copy stream, next to whatever the
    // compute stream still has queued (the substeps of the previous round)
    cudaEvent_t ev_ex_ready = nullptr, ev_ex_free = nullptr;
    bool ex_in_use = false;
};

namespace eut {

// ---- hi+lo accumulation -----------------------------------------------------------------------
struct dd { double hi, lo; };
__device__ __forceinline__ void dd_add(dd &s, double x)
{
    const double t = __dadd_rn(s.hi, x);
    const double bb = __dsub_rn(t, s.hi);
    const double err = __dadd_rn(__dsub_rn(s.hi, __dsub_rn(t, bb)), __dsub_rn(x, bb));
    s.hi = t;
    s.lo = __dadd_rn(s.lo, err);
}
__device__ __forceinline__ double dd_value(const dd &s)
{
    // non-finite sums: TwoSum's error term is NaN there, the plain sum is what R would return
    return (isfinite(s.hi) && isfinite(s.lo)) ? __dadd_rn(s.hi, s.lo) : s.hi;
}
__device__ __forceinline__ dd dd_warp_sum(dd s)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double h = __shfl_xor_sync(0xffffffffu, s.hi, o);
        const double l = __shfl_xor_sync(0xffffffffu, s.lo, o);
        dd_add(s, h);
        s.lo = __dadd_rn(s.lo, l);
    }
    return s;
}

template <typename T>
static int grow(T **p, int64_t *cap, int64_t need)
{
    if (need <= *cap && *p) return EUT_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    int64_t n = std::max<int64_t>(need + need / 8, 16);
    EUT_CUDA(cudaMalloc((void **)p, (size_t)n * sizeof(T)));
    *cap = n;
    return EUT_OK;
}

static int ensure_tmp(eut_cells *c, size_t bytes)
{
    if (bytes <= c->tmp_bytes && c->d_tmp) return EUT_OK;
    if (c->d_tmp) cudaFree(c->d_tmp);
    c->d_tmp = nullptr;
    EUT_CUDA(cudaMalloc(&c->d_tmp, bytes + bytes / 8 + 256));
    c->tmp_bytes = bytes + bytes / 8 + 256;
    return EUT_OK;
}

static int alloc_entries(eut_cells *c, int64_t M, int64_t T)
{
    if (M + 1 > c->cap_M || !c->d_cell_ptr) {
        cudaFree(c->d_cell_ptr); cudaFree(c->d_dmax); cudaFree(c->d_fsum);
        int64_t n = M + 1 + M / 8 + 16;
        EUT_CUDA(cudaMalloc((void **)&c->d_cell_ptr, n * sizeof(int64_t)));
        EUT_CUDA(cudaMalloc((void **)&c->d_dmax, n * sizeof(double)));
        EUT_CUDA(cudaMalloc((void **)&c->d_fsum, n * sizeof(double)));
        c->cap_M = n;
    }
    if (T > c->cap_T || !c->d_fid) {
        cudaFree(c->d_fid); cudaFree(c->d_fq); cudaFree(c->d_dist); cudaFree(c->d_acc);
        cudaFree(c->d_ecell); cudaFree(c->d_t_entry); cudaFree(c->d_t_key);
        cudaFree(c->d_seg_start); cudaFree(c->d_seg_fq);
        int64_t n = T + T / 8 + 16;
        EUT_CUDA(cudaMalloc((void **)&c->d_fid, n * sizeof(int32_t)));
        EUT_CUDA(cudaMalloc((void **)&c->d_fq, n * sizeof(int32_t)));
        EUT_CUDA(cudaMalloc((void **)&c->d_dist, n * sizeof(double)));
        EUT_CUDA(cudaMalloc((void **)&c->d_acc, n * sizeof(double)));
        EUT_CUDA(cudaMalloc((void **)&c->d_ecell, n * sizeof(int32_t)));
        EUT_CUDA(cudaMalloc((void **)&c->d_t_entry, n * sizeof(int32_t)));
        EUT_CUDA(cudaMalloc((void **)&c->d_t_key, n * sizeof(int32_t)));
        EUT_CUDA(cudaMalloc((void **)&c->d_seg_start, (n + 1) * sizeof(int32_t)));
        EUT_CUDA(cudaMalloc((void **)&c->d_seg_fq, n * sizeof(int32_t)));
        c->cap_T = n;
    }
    c->M = M;
    c->T = T;
    c->transposed = false;
    c->g_sorted = false;
    c->t_packed = false;
    c->fmol_field = nullptr;
    return EUT_OK;
}

// ---- per-entry / per-cell derived data --------------------------------------------------------
__global__ void __launch_bounds__(256)
k_entry_cell(const int64_t *__restrict__ cell_ptr, int32_t *__restrict__ ecell, const int n_cells)
{
    // one warp per cell
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_cells) return;
    for (int64_t k = cell_ptr[w] + lane; k < cell_ptr[w + 1]; k += 32) ecell[k] = w;
}

__global__ void __launch_bounds__(256)
k_entry_pos(const int32_t *__restrict__ fid, const int32_t *__restrict__ perm,
            int32_t *__restrict__ fq, const int64_t n_entries, const int n_points, int *bad)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_entries) return;
    const int f = fid[k];
    if (f < 1 || f > n_points) { atomicExch(bad, 1); fq[k] = 0; return; }
    fq[k] = perm[f - 1];
}

// field.frac ingredients per cell (R/run_simulation.R:729-730): max distance, sum(max - dist)
__global__ void __launch_bounds__(256)
k_cell_frac(const int64_t *__restrict__ cell_ptr, const double *__restrict__ dist,
            double *__restrict__ dmax, double *__restrict__ fsum, const int n_cells)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_cells) return;
    const int64_t k0 = cell_ptr[w], k1 = cell_ptr[w + 1];
    double m = -INFINITY;
    bool nan = false;
    for (int64_t k = k0 + lane; k < k1; k += 32) { double d = dist[k]; nan |= (d != d); m = fmax(m, d); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        nan |= (bool)__shfl_xor_sync(0xffffffffu, (int)nan, o);
    }
    dd s{0.0, 0.0};
    for (int64_t k = k0 + lane; k < k1; k += 32) dd_add(s, __dsub_rn(m, dist[k]));
    s = dd_warp_sum(s);
    if (lane == 0) { dmax[w] = m; fsum[w] = dd_value(s); }
    (void)nan;
}

// transposed list boundaries
__global__ void __launch_bounds__(256)
k_seg_flags(const int32_t *__restrict__ key, int32_t *__restrict__ flag, const int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    flag[k] = (k == 0 || key[k] != key[k - 1]) ? 1 : 0;
}
__global__ void __launch_bounds__(256)
k_seg_fill(const int32_t *__restrict__ key, const int32_t *__restrict__ flag,
           const int32_t *__restrict__ rank, int32_t *__restrict__ seg_start,
           int32_t *__restrict__ seg_fq, const int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (flag[k]) { seg_start[rank[k] - 1] = (int32_t)k; seg_fq[rank[k] - 1] = key[k]; }
}

static int build_transposed(eut_cells *c, cudaStream_t s)
{
    if (c->transposed) return EUT_OK;
    const int64_t T = c->T;
    c->n_seg = 0;
    if (T == 0) { c->transposed = true; return EUT_OK; }
    // stable LSD radix sort of (field position, entry index): entries of one field keep cell order
    size_t sort_bytes = 0;
    int32_t *vals_in = nullptr;
    int bits = 1;
    while ((1ll << bits) < c->plan->Np) bits++;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, c->d_fq, c->d_t_key, vals_in, c->d_t_entry,
                                    (int)T, 0, bits, s);
    size_t scan_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, scan_bytes, (int32_t *)nullptr, (int32_t *)nullptr, (int)T, s);
    const size_t iota_off = ((sort_bytes + 255) / 256) * 256;
    const size_t need = iota_off + std::max<size_t>((size_t)T * 4 * 2, scan_bytes + (size_t)T * 8 + 512);
    EUT_TRY(ensure_tmp(c, need + (size_t)T * 8 + 1024));
    char *base = (char *)c->d_tmp;
    vals_in = (int32_t *)(base + iota_off);
    {
        std::vector<int32_t> iota(T);
        std::iota(iota.begin(), iota.end(), 0);
        EUT_CUDA(cudaMemcpyAsync(vals_in, iota.data(), T * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        EUT_CUDA(cudaStreamSynchronize(s));
    }
    EUT_CUDA(cub::DeviceRadixSort::SortPairs(base, sort_bytes, c->d_fq, c->d_t_key, vals_in, c->d_t_entry,
                                             (int)T, 0, bits, s));
    int32_t *flag = (int32_t *)(base + iota_off);
    int32_t *rank = flag + T;
    void *scan_tmp = (void *)(((uintptr_t)(rank + T) + 255) / 256 * 256);
    const unsigned nb = (unsigned)((T + 255) / 256);
    k_seg_flags<<<nb, 256, 0, s>>>(c->d_t_key, flag, T);
    EUT_CUDA(cub::DeviceScan::InclusiveSum(scan_tmp, scan_bytes, flag, rank, (int)T, s));
    k_seg_fill<<<nb, 256, 0, s>>>(c->d_t_key, flag, rank, c->d_seg_start, c->d_seg_fq, T);
    int32_t n_seg = 0;
    EUT_CUDA(cudaMemcpyAsync(&n_seg, rank + (T - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    EUT_CUDA(cudaStreamSynchronize(s));
    c->n_seg = n_seg;
    const int32_t t32 = (int32_t)T;
    EUT_CUDA(cudaMemcpyAsync(c->d_seg_start + n_seg, &t32, sizeof(int32_t), cudaMemcpyHostToDevice, s));
    EUT_CUDA(cudaStreamSynchronize(s));
    c->launches += 4;
    c->transposed = true;
    return EUT_OK;
}

static int finish_lists(eut_cells *c, cudaStream_t s)
{
    // internal positions, entry->cell, field.frac ingredients
    const eut_plan *p = c->plan;
    if (c->T > 0) {
        int *bad = nullptr;
        EUT_TRY(ensure_tmp(c, 256));
        bad = (int *)c->d_tmp;
        EUT_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
        k_entry_pos<<<(unsigned)((c->T + 255) / 256), 256, 0, s>>>(c->d_fid, p->d_perm, c->d_fq, c->T, (int)p->N, bad);
        int h_bad = 0;
        EUT_CUDA(cudaMemcpyAsync(&h_bad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
        EUT_CUDA(cudaStreamSynchronize(s));
        if (h_bad) { set_error("index out of bounds: a cell's field.id is outside 1..%lld", (long long)p->N); return EUT_ERR_INDEX; }
    }
    if (c->M > 0) {
        const unsigned nb = (unsigned)((c->M * 32 + 255) / 256);
        k_entry_cell<<<nb, 256, 0, s>>>(c->d_cell_ptr, c->d_ecell, (int)c->M);
        k_cell_frac<<<nb, 256, 0, s>>>(c->d_cell_ptr, c->d_dist, c->d_dmax, c->d_fsum, (int)c->M);
    }
    c->launches += 3;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

// ---- gather ------------------------------------------------------------------------------------
// One warp per cell, cells in the order of their position in the field.  Lanes stride over the
// cell's entries in the order of their internal field position (g_idx: sorted once per cell map), so
// the 32 loads of a warp request fall on a few runs of consecutive addresses (x-neighbours of one
// raster row) instead of 32 different sectors.  A lane loads its (position, acc.prop) pairs ONCE and
// reuses them for all compound columns (G at a time, all loads independent); warps that run together
// belong to neighbouring cells, whose field points overlap, so the columns are read from HBM about
// once (the first version walked column groups in the outer loop and re-read the entry lists and the
// field 8x: 10 GB of DRAM reads per call, profiles/r01_cells_ncu.csv).  Sums are hi+lo pairs, so the
// walking order does not show in the rounded result.
template <int G, int EPL>
__global__ void __launch_bounds__(256)
k_gather(const double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
         const int32_t *__restrict__ cur, const int64_t *__restrict__ cell_ptr,
         const int32_t *__restrict__ g_idx, const int32_t *__restrict__ fq,
         const double *__restrict__ acc, const double field_vol, double *__restrict__ fmol,
         const int n_cells, const int n_cols, const int32_t *__restrict__ cell_order)
{
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (int64_t)n_cells) return;
    const int i = cell_order[(int)w];
    const int64_t k0 = cell_ptr[i], k1 = cell_ptr[i + 1];
    const bool one_pass = (k1 - k0) <= 32 * EPL;           // the usual case: the whole list sits in registers
    int q[EPL];
    double a[EPL];
    auto load_entries = [&](int64_t kb) {
#pragma unroll
        for (int m = 0; m < EPL; m++) {
            const int64_t k = kb + lane + 32 * m;
            q[m] = -1; a[m] = 0.0;
            if (k < k1) { const int e = g_idx[k]; q[m] = fq[e]; a[m] = acc[e]; }
        }
    };
    if (one_pass) load_entries(k0);
    for (int c0 = 0; c0 < n_cols; c0 += G) {
        dd s[G];
#pragma unroll
        for (int u = 0; u < G; u++) s[u] = dd{0.0, 0.0};
        for (int64_t kb = k0; kb < k1; kb += 32 * EPL) {
            if (!one_pass) load_entries(kb);
#pragma unroll
            for (int u = 0; u < G; u++) {
                const int c = min(c0 + u, n_cols - 1);
                const double *__restrict__ col = buf + (int64_t)cur[c] * buf_stride + (int64_t)c * col_stride;
                double v[EPL];
#pragma unroll
                for (int m = 0; m < EPL; m++) v[m] = q[m] >= 0 ? __ldg(col + q[m]) : 0.0;
                // sum_k ((conc_k / 1000) * fieldVol) * acc_k  (R/scavenge_compounds.R:17-22) is evaluated
                // as ((sum_k conc_k * acc_k) / 1000) * fieldVol: all terms are non-negative, the two
                // forms differ by a few ulp (<< the 1e-12 bar; R's own long double sum is not
                // reproducible to the last bit either) and the per-term division leaves the hot loop.
#pragma unroll
                for (int m = 0; m < EPL; m++)
                    if (q[m] >= 0) dd_add(s[u], __dmul_rn(v[m], a[m]));
            }
        }
#pragma unroll
        for (int u = 0; u < G; u++) {
            const dd r = dd_warp_sum(s[u]);
            if (lane == 0 && c0 + u < n_cols)
                fmol[(int64_t)(c0 + u) * n_cells + i] = __dmul_rn(__ddiv_rn(dd_value(r), 1000.0), field_vol);
        }
    }
}

// ---- gather, second generation ------------------------------------------------------------------
// The field is compound-major (one column = one contiguous vector: what the stencil and the column
// transfers want), so a cell's ~133 points of one column are ~21 short runs and every 8-byte load of
// the kernel above pulls a 32-byte sector: 366 M sectors from L2 per call, which is what bounds it.
// Here the field is first copied point-major (one pass over HBM, 0.16 ms at 1 M x 64), then a warp
// takes a cell and its lanes take the COLUMNS: the 64 values of one field point are 512 contiguous
// bytes, every sector is used in full, and neighbouring cells hit the same lines in L1 / L2.  Each
// lane adds its column's terms in the cell's position order as a hi+lo pair: no warp reduction, and
// the result does not depend on the launch shape.
__global__ void __launch_bounds__(256)
k_field_transpose(const double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
                  const int32_t *__restrict__ cur, double *__restrict__ tf, const int n_padded,
                  const int n_cols, const int ldc)
{
    __shared__ double tile[32][33];
    const int q0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;          // 32 x 8
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int c = c0 + j, q = q0 + tx;
        double v = 0.0;
        if (c < n_cols && q < n_padded) v = buf[(int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + q];
        tile[j][tx] = v;
    }
    __syncthreads();
#pragma unroll
    for (int j = ty; j < 32; j += 8) {
        const int q = q0 + j, c = c0 + tx;
        if (q < n_padded && c < ldc) tf[(int64_t)q * ldc + c] = tile[tx][j];
    }
}

template <int CPL>
__global__ void __launch_bounds__(256)
k_gather_t(const double *__restrict__ tf, const int ldc, const int64_t *__restrict__ cell_ptr,
           const int32_t *__restrict__ g_idx, const int32_t *__restrict__ fq, const double *__restrict__ acc,
           const double field_vol, double *__restrict__ fmol, const int n_cells, const int n_cols,
           const int32_t *__restrict__ cell_order)
{
    const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (int64_t)n_cells) return;
    const unsigned full = 0xffffffffu;
    const int i = cell_order[(int)w];
    const int64_t k0 = cell_ptr[i], k1 = cell_ptr[i + 1];
    for (int c0 = 0; c0 < n_cols; c0 += 32 * CPL) {
        dd s[CPL];
#pragma unroll
        for (int u = 0; u < CPL; u++) s[u] = dd{0.0, 0.0};
        for (int64_t kb = k0; kb < k1; kb += 32) {
            int ql = 0;
            double al = 0.0;
            if (kb + lane < k1) { const int e = g_idx[kb + lane]; ql = fq[e]; al = acc[e]; }
            const int nr = (int)min((int64_t)32, k1 - kb);
#pragma unroll 4
            for (int r = 0; r < nr; r++) {
                const int q = __shfl_sync(full, ql, r);
                const double a = __shfl_sync(full, al, r);
                const double *__restrict__ row = tf + (int64_t)q * ldc + c0 + lane;
#pragma unroll
                for (int u = 0; u < CPL; u++)
                    if (c0 + lane + 32 * u < n_cols) dd_add(s[u], __dmul_rn(__ldg(row + 32 * u), a));
            }
        }
#pragma unroll
        for (int u = 0; u < CPL; u++) {
            const int c = c0 + lane + 32 * u;
            // ((sum_k conc_k * acc_k) / 1000) * fieldVol, see k_gather
            if (c < n_cols) fmol[(int64_t)c * n_cells + i] = __dmul_rn(__ddiv_rn(dd_value(s[u]), 1000.0), field_vol);
        }
    }
}

// key of the per-cell position sort: (cell << 32) | internal position
__global__ void __launch_bounds__(256)
k_gather_keys(const int32_t *__restrict__ ecell, const int32_t *__restrict__ fq,
              unsigned long long *__restrict__ keys, int32_t *__restrict__ vals, const int64_t n)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    keys[k] = ((unsigned long long)(uint32_t)ecell[k] << 32) | (uint32_t)fq[k];
    vals[k] = (int32_t)k;
}

__global__ void __launch_bounds__(256)
k_cell_keys(const int64_t *__restrict__ cell_ptr, const int32_t *__restrict__ g_idx,
            const int32_t *__restrict__ fq, int32_t *__restrict__ keys, int32_t *__restrict__ vals,
            const int n_cells)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    keys[i] = cell_ptr[i + 1] > cell_ptr[i] ? fq[g_idx[cell_ptr[i]]] : 0x7fffffff;
    vals[i] = i;
}

static int build_gather_order(eut_cells *c, cudaStream_t s)
{
    if (c->g_sorted) return EUT_OK;
    const int64_t T = c->T;
    if (T > 0) {
        EUT_TRY(grow(&c->d_g_idx, &c->g_cap, T));
        int bits = 1;
        while ((1ll << bits) < c->plan->Np) bits++;
        int cbits = 1;
        while ((1ll << cbits) < c->M + 1) cbits++;
        size_t sort_bytes = 0;
        unsigned long long *keys_in = nullptr, *keys_out = nullptr;
        int32_t *vals_in = nullptr;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, keys_in, keys_out, vals_in, c->d_g_idx, (int)T, 0,
                                        32 + cbits, s);
        const size_t off_keys = ((sort_bytes + 255) / 256) * 256;
        EUT_TRY(ensure_tmp(c, off_keys + (size_t)T * (8 + 8 + 4) + 1024));
        char *base = (char *)c->d_tmp;
        keys_in = (unsigned long long *)(base + off_keys);
        keys_out = keys_in + T;
        vals_in = (int32_t *)(keys_out + T);
        k_gather_keys<<<(unsigned)((T + 255) / 256), 256, 0, s>>>(c->d_ecell, c->d_fq, keys_in, vals_in, T);
        // position bits below bit `bits` and cell bits from bit 32 up: sort both ranges in one go
        EUT_CUDA(cub::DeviceRadixSort::SortPairs(base, sort_bytes, keys_in, keys_out, vals_in, c->d_g_idx,
                                                 (int)T, 0, 32 + cbits, s));
        (void)bits;
        c->launches += 2;
        EUT_CUDA(cudaGetLastError());
    }
    // cells ordered by the field position of their first (lowest) entry
    if (c->M > 0) {
        EUT_TRY(grow(&c->d_cell_order, &c->order_cap, c->M));
        size_t sort_bytes = 0;
        int32_t *k_in = nullptr, *k_out = nullptr, *v_in = nullptr;
        cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, k_in, k_out, v_in, c->d_cell_order, (int)c->M, 0, 32, s);
        const size_t off_k = ((sort_bytes + 255) / 256) * 256;
        EUT_TRY(ensure_tmp(c, off_k + (size_t)c->M * 12 + 1024));
        char *base = (char *)c->d_tmp;
        k_in = (int32_t *)(base + off_k);
        k_out = k_in + c->M;
        v_in = k_out + c->M;
        k_cell_keys<<<(unsigned)((c->M + 255) / 256), 256, 0, s>>>(c->d_cell_ptr, c->d_g_idx, c->d_fq, k_in, v_in, (int)c->M);
        EUT_CUDA(cub::DeviceRadixSort::SortPairs(base, sort_bytes, k_in, k_out, v_in, c->d_cell_order, (int)c->M, 0, 32, s));
        c->launches += 2;
        EUT_CUDA(cudaGetLastError());
    }
    c->g_sorted = true;
    return EUT_OK;
}

// ---- scatter -----------------------------------------------------------------------------------
// colSums(accmat) for every uptake exchange (R/run_simulation.R:696): one warp per exchange entry
__global__ void __launch_bounds__(256)
k_ex_colsum(const double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
            const int32_t *__restrict__ cur, const int64_t *__restrict__ cell_ptr,
            const int32_t *__restrict__ fq, const double *__restrict__ acc, const double field_vol,
            const int64_t *__restrict__ ex_ptr, const int32_t *__restrict__ ex_cpd,
            const double *__restrict__ ex_flx, double *__restrict__ ex_cs, const int n_cells,
            const double vol_l)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_cells) return;
    for (int64_t e = ex_ptr[i]; e < ex_ptr[i + 1]; e++) {
        const double ex = ex_flx[e];
        if (!(ex < 0)) {
            // production: mM if all went to one field, (ex / 1e12) / (fieldVol / 1e15)  (:732)
            if (lane == 0) ex_cs[e] = __ddiv_rn(__ddiv_rn(ex, 1e12), vol_l);
            continue;
        }
        const int c = ex_cpd[e] - 1;
        const double *__restrict__ col = buf + (int64_t)cur[c] * buf_stride + (int64_t)c * col_stride;
        dd s{0.0, 0.0};
        for (int64_t k = cell_ptr[i] + lane; k < cell_ptr[i + 1]; k += 32)
            dd_add(s, __dmul_rn(col[fq[k]], acc[k]));   // same form as the gather
        s = dd_warp_sum(s);
        // The reference evaluates, per field, am = fmolPerField / colSum and Delta = am * ex / 1e12 /
        // (fieldVol / 1e15) (R/run_simulation.R:696, :707-708) with fmolPerField = conc / 1000 * fieldVol
        // * acc.prop.  Everything that does not depend on the field is folded into one factor per
        // exchange here, so that the scatter kernel needs two multiplications per contribution instead
        // of four divisions (a division is ~25 instructions; the scatter was bound by instruction
        // issue): Delta = conc * acc.prop * g,  g = ex / colSum / 1e12 / vol_l * fieldVol / 1000.
        // A few ulp from the reference's operation order, far inside 1e-12; colSum = 0 gives
        // g = -Inf and 0 * -Inf = NaN, the reference's 0 / 0.
        if (lane == 0) {
            const double cs = __dmul_rn(__ddiv_rn(dd_value(s), 1000.0), field_vol);
            const double g = __ddiv_rn(__ddiv_rn(__ddiv_rn(ex, cs), 1e12), vol_l);
            ex_cs[e] = __ddiv_rn(__dmul_rn(g, field_vol), 1000.0);
        }
    }
}

// One warp per touched field point.  The warp walks the (cell, entry) pairs of its field in cell
// order; the exchanges of one cell are evaluated in parallel (lane = exchange: compound, flux,
// colSum and -- for an uptake -- the field's own concentration are one coalesced / gathered load
// each) and added by that lane to the warp's per-column accumulators in shared memory.  A cell lists
// a compound at most once (checked on the host), so the lanes of one step hit different columns, and
// the steps are separated by __syncwarp: every (field, compound) sum is taken in cell order, like the
// reference's by-group sum over the concatenated per-cell tables (R/run_simulation.R:309-310); the
// reference's "I.up rows, then I.pd rows" order inside one cell cannot matter for the same reason.
// Finally: NaN -> 0, add, clamp (:315-319).  (The first version kept the accumulators in registers
// and routed every contribution to its owner lane by shuffle: 1 600 instructions per field point.)
template <int COLS>
__global__ void __launch_bounds__(256)
k_scatter(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
          const int32_t *__restrict__ cur, const int32_t *__restrict__ seg_start,
          const int32_t *__restrict__ seg_fq, const int32_t *__restrict__ t_entry,
          const int32_t *__restrict__ ecell, const double *__restrict__ dist,
          const double *__restrict__ acc, const double *__restrict__ dmax,
          const double *__restrict__ fsum, const double field_vol, const double vol_l,
          const int64_t *__restrict__ ex_ptr, const int32_t *__restrict__ ex_cpd,
          const double *__restrict__ ex_flx, const double *__restrict__ ex_cs,
          const uint8_t *__restrict__ is_const, const int n_seg, const int c0, const int n_cols)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n_seg) return;
    const unsigned full = 0xffffffffu;
    __shared__ double s_sum[8][COLS];
    __shared__ uint8_t s_hit[8][COLS];
    double *sum = s_sum[threadIdx.x >> 5];
    uint8_t *hit = s_hit[threadIdx.x >> 5];
#pragma unroll
    for (int u = lane; u < COLS; u += 32) { sum[u] = 0.0; hit[u] = 0; }
    __syncwarp();
    const int q = seg_fq[t];
    const int r_begin = seg_start[t], r_end = seg_start[t + 1];
    for (int rb = r_begin; rb < r_end; rb += 32) {
        // the per-cell quantities of up to 32 incident cells are fetched by the lanes in parallel (one
        // chain of dependent loads for all of them instead of one per cell), then handed out by shuffle
        double l_ak = 0.0, l_frac = 0.0;
        long long l_e0 = 0;
        int l_ne = 0;
        if (rb + lane < r_end) {
            const int k = t_entry[rb + lane];
            const int i = ecell[k];
            l_ak = acc[k];
            l_frac = __ddiv_rn(__dsub_rn(dmax[i], dist[k]), fsum[i]);   // :729-730
            l_e0 = ex_ptr[i];
            l_ne = (int)(ex_ptr[i + 1] - l_e0);
        }
        const int nr = min(32, r_end - rb);
        for (int rr = 0; rr < nr; rr++) {
            const double ak = __shfl_sync(full, l_ak, rr);
            const double frac = __shfl_sync(full, l_frac, rr);
            const int64_t e0 = __shfl_sync(full, l_e0, rr);
            const int ne = __shfl_sync(full, l_ne, rr);
            for (int base = 0; base < ne; base += 32) {
                if (base + lane < ne) {
                    const int64_t e = e0 + base + lane;
                    const int c = ex_cpd[e] - 1;
                    const double ex = ex_flx[e];
                    if (c >= c0 && c < c0 + COLS && c < n_cols && ex != 0.0 && ex == ex) {
                        double d;
                        if (ex < 0) {
                            const double cv = buf[(int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + q];
                            d = __dmul_rn(__dmul_rn(cv, ak), ex_cs[e]);                    // :696, :707-708 (see k_ex_colsum)
                        } else {
                            d = __dmul_rn(frac, ex_cs[e]);                                 // :732-734
                        }
                        sum[c - c0] = __dadd_rn(sum[c - c0], d);
                        hit[c - c0] = 1;
                    }
                }
                __syncwarp();
            }
        }
    }
#pragma unroll
    for (int u = lane; u < COLS; u += 32) {
        const int c = c0 + u;
        if (!hit[u] || c >= n_cols) continue;
        if (is_const && is_const[c]) continue;                          // :311-312
        double d = sum[u];
        if (d != d) d = 0.0;                                             // :315
        double *pc = buf + (int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + q;
        double v = __dadd_rn(*pc, d);                                    // :318
        if (v < 1e-12) v = 0.0;                                          // :319
        *pc = v;
    }
}

// ---- scatter, second generation --------------------------------------------------------------------
// Lanes = compound columns, like the second gather.  The sparse exchange list of every cell is first
// spread into a dense row of factors (k_ex_dense: 8 bytes per cell and column, L2 resident) with two
// bit masks per 64 columns (uptake / production); the warp of a field point then walks its incident
// cells in cell order and every lane adds what the cell does to ITS columns:
//     uptake      Delta = conc[q, c] * acc.prop * g          (g folded per exchange, see k_ex_colsum)
//     production  Delta = (dmax - dist) / fsum * ex'          (R/run_simulation.R:729-734)
// into a register.  Same terms, same (cell) order per (field, compound) sum as the first version --
// bit-identical results -- but no exchange-list walk, no shared-memory accumulators and a third of
// the instructions.  Finally NaN -> 0, add, clamp (:315-319).
__global__ void __launch_bounds__(256)
k_ex_dense(const int64_t *__restrict__ ex_ptr, const int32_t *__restrict__ ex_cpd, const double *__restrict__ ex_flx,
           const double *__restrict__ ex_cs, double *__restrict__ exd, unsigned long long *__restrict__ exmask,
           const int n_cells, const int n_cols, const int ldc)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_cells) return;
    for (int64_t e = ex_ptr[i] + lane; e < ex_ptr[i + 1]; e += 32) {
        const int c = ex_cpd[e] - 1;
        const double ex = ex_flx[e];
        if (c < 0 || c >= n_cols || ex == 0.0 || ex != ex) continue;
        exd[(int64_t)i * ldc + c] = ex_cs[e];
        atomicOr(exmask + ((int64_t)i * (ldc / 64) + c / 64) * 2 + (ex < 0 ? 0 : 1), 1ull << (c & 63));
    }
}

template <int CPL>
__global__ void __launch_bounds__(256)
k_scatter_t(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
            const int32_t *__restrict__ cur, const int32_t *__restrict__ seg_start,
            const int32_t *__restrict__ seg_fq, const int32_t *__restrict__ t_entry,
            const int32_t *__restrict__ ecell, const double *__restrict__ dist,
            const double *__restrict__ acc, const double *__restrict__ dmax,
            const double *__restrict__ fsum, const double *__restrict__ exd,
            const unsigned long long *__restrict__ exmask, const int ldc,
            const uint8_t *__restrict__ is_const, const int n_seg, const int n_cols)
{
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n_seg) return;
    const unsigned full = 0xffffffffu;
    const int q = seg_fq[t];
    const int r_begin = seg_start[t], r_end = seg_start[t + 1];
    for (int c0 = 0; c0 < n_cols; c0 += 64) {              // 64 columns per pass: one mask pair per cell
        double sum[2] = {0.0, 0.0}, cv[2] = {0.0, 0.0};
        double *pc[2];
        unsigned hit = 0;
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int c = min(c0 + lane + 32 * u, n_cols - 1);
            pc[u] = buf + (int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + q;
            cv[u] = *pc[u];
        }
        for (int rb = r_begin; rb < r_end; rb += 32) {
            // the per-cell quantities of up to 32 incident cells are fetched by the lanes in parallel
            double l_ak = 0.0, l_frac = 0.0;
            int l_i = 0;
            unsigned long long l_mu = 0, l_mp = 0;
            if (rb + lane < r_end) {
                const int k = t_entry[rb + lane];
                l_i = ecell[k];
                l_ak = acc[k];
                l_frac = __ddiv_rn(__dsub_rn(dmax[l_i], dist[k]), fsum[l_i]);   // :729-730
                const unsigned long long *m = exmask + ((int64_t)l_i * (ldc / 64) + c0 / 64) * 2;
                l_mu = m[0]; l_mp = m[1];
            }
            const int nr = min(32, r_end - rb);
            for (int rr = 0; rr < nr; rr++) {
                const unsigned long long mu = __shfl_sync(full, l_mu, rr), mp = __shfl_sync(full, l_mp, rr);
                if ((mu | mp) == 0ull) continue;                                  // warp-uniform
                const int i = __shfl_sync(full, l_i, rr);
                const double ak = __shfl_sync(full, l_ak, rr);
                const double frac = __shfl_sync(full, l_frac, rr);
                const double *__restrict__ row = exd + (int64_t)i * ldc + c0 + lane;
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int b = lane + 32 * u;
                    const bool up = (mu >> b) & 1ull, pd = (mp >> b) & 1ull;
                    if (up | pd) {
                        const double f = __ldg(row + 32 * u);
                        const double d = up ? __dmul_rn(__dmul_rn(cv[u], ak), f) : __dmul_rn(frac, f);
                        sum[u] = __dadd_rn(sum[u], d);
                        hit |= 1u << u;
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 2; u++) {
            const int c = c0 + lane + 32 * u;
            if (!((hit >> u) & 1u) || c >= n_cols) continue;
            if (is_const && is_const[c]) continue;                          // :311-312
            double d = sum[u];
            if (d != d) d = 0.0;                                             // :315
            double v = __dadd_rn(cv[u], d);                                  // :318
            if (v < 1e-12) v = 0.0;                                          // :319
            *pc[u] = v;
        }
    }
}

// ---- scatter, third generation ----------------------------------------------------------------------
// The second generation walks a field point's incident cells with lanes = columns: of the 64 column
// slots a cell visit spends instructions on, only the ~10 the cell exchanges do anything (583 M warp
// instructions, 12 % of the HBM rate: profiles/r01_cells_v3_ncu.csv), and every visit is a chain of
// dependent loads t_entry -> ecell -> mask -> factor row.
// Here FOUR LANES own a field point and a warp 8 consecutive touched points.  The warp first stages its
// 8 x COLS tile of the field in shared memory, then the lanes of a point walk ITS incident cells in
// cell order -- cell, acc.prop and the production share come
// packed in transposed order (k_t_pack, once per cell map) -- and, per cell, the cell's own compact
// exchange list: one 16-byte record per exchange (factor, column, kind), i.e. only the columns the cell
// exchanges.  Contributions are added into a per-point, per-column accumulator in shared memory
// (column-major, bank = lane: conflict free).  Same terms, same cell order per (field, compound) sum
// as before -- bit-identical results -- for a tenth of the instructions.  Finally NaN -> 0, add, clamp
// on the touched entries (R/run_simulation.R:315-319) and a coalesced write of the touched rows.
__global__ void __launch_bounds__(256)
k_t_pack(const int32_t *__restrict__ t_entry, const int32_t *__restrict__ ecell, const double *__restrict__ acc,
         const double *__restrict__ dist, const double *__restrict__ dmax, const double *__restrict__ fsum,
         int32_t *__restrict__ t_cell, double2 *__restrict__ t_af, const int64_t n)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    const int k = t_entry[r];
    const int i = ecell[k];
    t_cell[r] = i;
    t_af[r] = make_double2(acc[k], __ddiv_rn(__dsub_rn(dmax[i], dist[k]), fsum[i]));   // :729-730
}

// exchange records: factor (k_ex_colsum) | column (0-based), bit 30 = production; -1 = no contribution
__global__ void __launch_bounds__(256)
k_ex_pack(const int64_t *__restrict__ ex_ptr, const int32_t *__restrict__ ex_cpd, const double *__restrict__ ex_flx,
          const double *__restrict__ ex_cs, int32_t *__restrict__ ex_ptr32, double2 *__restrict__ exrec,
          const int n_cells, const int n_cols)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n_cells) return;
    ex_ptr32[i] = (int32_t)ex_ptr[i];
    if (i == n_cells) return;
    for (int64_t e = ex_ptr[i]; e < ex_ptr[i + 1]; e++) {
        const int c = ex_cpd[e] - 1;
        const double ex = ex_flx[e];
        long long code = -1;
        if (c >= 0 && c < n_cols && ex != 0.0 && ex == ex) code = (long long)c | (ex < 0 ? 0ll : (1ll << 30));
        exrec[e] = make_double2(ex_cs[e], __longlong_as_double(code));
    }
}

// L = 4 lanes share a field point (lane j takes the exchanges j, j + 4, ... of the visited cell: a cell
// lists a compound at most once, so the lanes of a point never meet in a column), a warp owns 8 touched
// points and a CTA of 4 warps 32: 8 KB of shared memory per warp, 24 warps per SM.
template <int COLS>
__global__ void __launch_bounds__(128)
k_scatter_p(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
            const int32_t *__restrict__ cur, const int32_t *__restrict__ seg_start,
            const int32_t *__restrict__ seg_fq, const int32_t *__restrict__ t_cell,
            const double2 *__restrict__ t_af, const int32_t *__restrict__ ex_ptr32,
            const double2 *__restrict__ exrec, const uint8_t *__restrict__ is_const, const int n_seg,
            const int c0, const int n_cols)
{
    constexpr int L = 4, P = 8, G = 3;                       // lanes per point, points per warp, exchanges in flight per lane
    constexpr int ROWS = COLS + 1;                           // row COLS: dummy column for exchanges this pass skips
    extern __shared__ __align__(16) double s_tile[];         // per warp [2][ROWS][P]: field values, sums
    __shared__ long long s_base[COLS];                       // offset of column c0 + c inside buf (current buffer)
    const unsigned full = 0xffffffffu;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = lane >> 2, j = lane & 3;
    double *cv = s_tile + (size_t)warp * 2 * ROWS * P, *sum = cv + ROWS * P;
    const int nc = min(COLS, n_cols - c0);
    for (int c = threadIdx.x; c < COLS; c += blockDim.x)
        s_base[c] = c < nc ? (long long)cur[c0 + c] * buf_stride + (long long)(c0 + c) * col_stride : 0ll;
    __syncthreads();
    const int t = (blockIdx.x * 4 + warp) * P + p;
    const bool valid = t < n_seg;
    const int q = valid ? seg_fq[t] : 0;
    // tile cell of (column c, point p): the four lanes of a point work on different columns at the same
    // time; the xor spreads them over the banks (c * 8 + p alone puts all even columns of a point on one bank pair)
    auto cell = [](int c, int pp) { return c * P + (pp ^ ((c >> 1) & 7)); };
    // tile: lane (p, j) loads the columns j, j + 4, ...
#pragma unroll 4
    for (int c = j; c < ROWS; c += L) {
        double v = 0.0;
        if (valid && c < nc) v = buf[s_base[c] + q];
        cv[cell(c, p)] = v;
        sum[cell(c, p)] = 0.0;
    }
    unsigned long long hit = 0ull;
    const int r0 = valid ? seg_start[t] : 0, n_vis = valid ? seg_start[t + 1] - r0 : 0;
    const int max_vis = __reduce_max_sync(full, n_vis);
    // Software pipeline over the visits (the loads of a visit form a chain cell -> list bounds -> exchange
    // records, and a warp issues in order): iteration v works on visit v with everything in registers, and
    // issues the records of v + 1, the bounds of v + 2, the cell and weights of v + 3.
    const double2 none = make_double2(0.0, __longlong_as_double(-1ll));
    int i3 = 0, e0_2 = 0, e1_2 = 0, e0_1 = 0, e1_1 = 0;
    double2 af3 = make_double2(0.0, 0.0), af2 = af3, af1 = af3, rec1[G];
#pragma unroll
    for (int u = 0; u < G; u++) rec1[u] = none;
    auto stage1 = [&](int v) { if (v < n_vis) { i3 = __ldg(t_cell + r0 + v); af3 = __ldg(t_af + r0 + v); } };
    auto stage2 = [&](int v) { if (v < n_vis) { e0_2 = __ldg(ex_ptr32 + i3); e1_2 = __ldg(ex_ptr32 + i3 + 1); af2 = af3; } };
    auto stage3 = [&](int v) {
        e0_1 = e0_2; e1_1 = (v < n_vis) ? e1_2 : e0_2; af1 = af2;
#pragma unroll
        for (int u = 0; u < G; u++) rec1[u] = (e0_1 + j + u * L < e1_1) ? __ldg(exrec + e0_1 + j + u * L) : none;
    };
    stage1(0); stage2(0); stage1(1); stage3(0); stage2(1); stage1(2);
    __syncwarp();
    for (int v = 0; v < max_vis; v++) {
        const double2 af = af1;
        const int e0 = e0_1, e1 = e1_1;
        double2 rec[G];
#pragma unroll
        for (int u = 0; u < G; u++) rec[u] = rec1[u];
        stage3(v + 1); stage2(v + 2); stage1(v + 3);
        for (int eb = e0 + j; eb < e1; eb += L * G) {
            if (eb != e0 + j) {                              // lists longer than L * G exchanges: not pipelined
#pragma unroll
                for (int u = 0; u < G; u++) rec[u] = (eb + u * L < e1) ? __ldg(exrec + eb + u * L) : none;
            }
            // a cell lists a compound at most once (checked on the host): the G accumulators of one group, and
            // those of the other lanes of the point, are different cells of the tile -- independent chains
            int a[G];
            double sv[G], cvv[G];
            bool prod[G];
#pragma unroll
            for (int u = 0; u < G; u++) {
                const long long code = __double_as_longlong(rec[u].y);
                int c = (int)(code & 0x3fffffffll) - c0;
                const bool ok = code >= 0 && (unsigned)c < (unsigned)nc;
                if (!ok) c = COLS;
                prod[u] = (code & (1ll << 30)) != 0;
                a[u] = cell(c, p);
                if (ok) hit |= 1ull << c;
            }
#pragma unroll
            for (int u = 0; u < G; u++) { sv[u] = sum[a[u]]; cvv[u] = cv[a[u]]; }
#pragma unroll
            for (int u = 0; u < G; u++) {
                const double d = prod[u] ? __dmul_rn(af.y, rec[u].x)                          // production, :732-734
                                         : __dmul_rn(__dmul_rn(cvv[u], af.x), rec[u].x);      // uptake, :696, :707-708
                if (a[u] < COLS * P) sum[a[u]] = __dadd_rn(sv[u], d);
            }
        }
        __syncwarp();                                        // the next cell may touch a column another lane just wrote
    }
    hit |= __shfl_xor_sync(full, hit, 1);
    hit |= __shfl_xor_sync(full, hit, 2);
    if (!valid) return;
    for (int c = j; c < nc; c += L) {
        if (!((hit >> c) & 1ull)) continue;
        if (is_const && is_const[c0 + c]) continue;                      // :311-312
        double d = sum[cell(c, p)];
        if (d != d) d = 0.0;                                             // :315
        double v = __dadd_rn(cv[cell(c, p)], d);                         // :318
        if (v < 1e-12) v = 0.0;                                          // :319
        buf[s_base[c] + q] = v;
    }
}

// colSums from the gather that ran on this very state of the field: g per exchange without touching the field
__global__ void __launch_bounds__(256)
k_ex_factor(const double *__restrict__ fmol, const int64_t *__restrict__ ex_ptr, const int32_t *__restrict__ ex_cpd,
            const double *__restrict__ ex_flx, double *__restrict__ ex_cs, const int n_cells, const double field_vol,
            const double vol_l)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    for (int64_t e = ex_ptr[i]; e < ex_ptr[i + 1]; e++) {
        const double ex = ex_flx[e];
        if (!(ex < 0)) { ex_cs[e] = __ddiv_rn(__ddiv_rn(ex, 1e12), vol_l); continue; }
        const double cs = fmol[(int64_t)(ex_cpd[e] - 1) * n_cells + i];
        const double g = __ddiv_rn(__ddiv_rn(__ddiv_rn(ex, cs), 1e12), vol_l);
        ex_cs[e] = __ddiv_rn(__dmul_rn(g, field_vol), 1000.0);
    }
}

// ---- chemotaxis moments (R/run_simulation.R:260-291) -------------------------------------------------
// Per cell and chemotaxis compound the reference takes, over the cell's field list, the weighted mean of
// the normalised offsets (dx, dy) = (field.xy - cell.xy) / (size/2 + scavengeDist) with weights
// conc / sum(conc) (uniform when the sum is 0), and the weighted mean of conc itself.  The kernel
// returns the sums those means are made of -- sum c, sum c^2, sum dx c, sum dy c, sum dx, sum dy, n --
// as hi+lo accumulations (one warp per cell); the Hill term and the scaling stay with the caller.
__global__ void __launch_bounds__(256)
k_chemo_moments(const double *__restrict__ col, const double *__restrict__ px, const double *__restrict__ py,
                const int64_t *__restrict__ cell_ptr, const int32_t *__restrict__ fq, const double *__restrict__ cxy,
                double *__restrict__ out, const int n_cells)
{
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= n_cells) return;
    const double cx = cxy[i], cy = cxy[n_cells + i], rr = cxy[2 * (int64_t)n_cells + i];
    dd s0{0, 0}, s1{0, 0}, sx{0, 0}, sy{0, 0}, ux{0, 0}, uy{0, 0};
    for (int64_t k = cell_ptr[i] + lane; k < cell_ptr[i + 1]; k += 32) {
        const int q = fq[k];
        const double c = col[q];
        const double dx = __ddiv_rn(__dsub_rn(px[q], cx), rr), dy = __ddiv_rn(__dsub_rn(py[q], cy), rr);   // :268-272
        dd_add(s0, c); dd_add(s1, __dmul_rn(c, c));
        dd_add(sx, __dmul_rn(dx, c)); dd_add(sy, __dmul_rn(dy, c));
        dd_add(ux, dx); dd_add(uy, dy);
    }
    s0 = dd_warp_sum(s0); s1 = dd_warp_sum(s1); sx = dd_warp_sum(sx); sy = dd_warp_sum(sy);
    ux = dd_warp_sum(ux); uy = dd_warp_sum(uy);
    if (lane == 0) {
        const int64_t M = n_cells;
        out[i] = dd_value(s0); out[M + i] = dd_value(s1); out[2 * M + i] = dd_value(sx); out[3 * M + i] = dd_value(sy);
        out[4 * M + i] = dd_value(ux); out[5 * M + i] = dd_value(uy); out[6 * M + i] = (double)(cell_ptr[i + 1] - cell_ptr[i]);
    }
}

// ---- claim rescale (R/run_simulation.R:236-241) ------------------------------------------------
__global__ void __launch_bounds__(256)
k_claim(const int32_t *__restrict__ seg_start, const int32_t *__restrict__ t_entry,
        double *__restrict__ acc, const int n_seg)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_seg) return;
    const int r0 = seg_start[t], r1 = seg_start[t + 1];
    dd s{0.0, 0.0};
    for (int r = r0; r < r1; r++) dd_add(s, acc[t_entry[r]]);
    const double claim = dd_value(s);
    if (!(claim > 1)) return;
    const double e = 2.718281828459045;  // exp(1)
    dd ts{0.0, 0.0};
    for (int r = r0; r < r1; r++) dd_add(ts, pow(acc[t_entry[r]], e));
    const double tsum = dd_value(ts);
    for (int r = r0; r < r1; r++) {
        const int k = t_entry[r];
        double tmp = pow(acc[k], e);
        tmp = __dmul_rn(__ddiv_rn(tmp, tsum), claim);
        acc[k] = __ddiv_rn(tmp, claim);
    }
}

// ---- lists from geometry (R/run_simulation.R:217-230 box prefilter, :591-624) -------------------
// Field points are pre-sorted by (z, x, y) (the gridDT key, :168) and grouped into runs of equal
// (z, x).  A thread handles one cell: for every layer whose z passes the box test, binary-search
// the runs whose x passes, and inside each run the points whose y passes -- with the very same
// floating-point predicates as the reference (|v - c| <= R is monotone in v) -- then apply the
// distance test.  Pass 0 counts, pass 1 writes, so the output order is the key order.
__device__ __forceinline__ bool in_box(double v, double c, double R) { return fabs(__dsub_rn(v, c)) <= R; }

template <bool WRITE>
__global__ void __launch_bounds__(128)
k_cellmap(const double *__restrict__ gx, const double *__restrict__ gy,
          const int32_t *__restrict__ gfield, const int32_t *__restrict__ col_start,
          const double *__restrict__ col_x, const int32_t *__restrict__ layer_col,
          const double *__restrict__ layer_z, const int n_layers,
          const double *__restrict__ cx, const double *__restrict__ cy,
          const double *__restrict__ csize, const double *__restrict__ cscv, const int n_cells,
          const int dist_mode, int64_t *__restrict__ counts, const int64_t *__restrict__ cell_ptr,
          int32_t *__restrict__ fid, double *__restrict__ fdist, double *__restrict__ facc)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const double x0 = cx[i], y0 = cy[i];
    const double R = __dadd_rn(__ddiv_rn(csize[i], 2.0), cscv[i]);
    const double inv = __ddiv_rn(-1.0, cscv[i]);
    int64_t n = 0;
    int64_t o = WRITE ? cell_ptr[i] : 0;
    for (int l = 0; l < n_layers; l++) {
        const double z = layer_z[l];
        if (!in_box(z, 0.0, R)) continue;
        // runs [a, b) of this layer with |x - x0| <= R
        int lo = layer_col[l], hi = layer_col[l + 1];
        int a = lo, b = hi;
        {   // first run with x - x0 >= -R  (predicate monotone in x)
            int l2 = lo, h2 = hi;
            while (l2 < h2) { int m = (l2 + h2) >> 1; if (__dsub_rn(col_x[m], x0) < -R) l2 = m + 1; else h2 = m; }
            a = l2;
            l2 = a; h2 = hi;   // first run with x - x0 > R
            while (l2 < h2) { int m = (l2 + h2) >> 1; if (__dsub_rn(col_x[m], x0) <= R) l2 = m + 1; else h2 = m; }
            b = l2;
        }
        for (int r = a; r < b; r++) {
            if (!in_box(col_x[r], x0, R)) continue;  // NaN safety; always true otherwise
            int p0 = col_start[r], p1 = col_start[r + 1];
            int l2 = p0, h2 = p1;
            while (l2 < h2) { int m = (l2 + h2) >> 1; if (__dsub_rn(gy[m], y0) < -R) l2 = m + 1; else h2 = m; }
            for (int pidx = l2; pidx < p1; pidx++) {
                const double dy = __dsub_rn(gy[pidx], y0);
                if (dy > R) break;
                if (!(fabs(dy) <= R)) continue;
                const double dx = __dsub_rn(gx[pidx], x0);
                double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
                if (dist_mode == 1) d2 = __dadd_rn(d2, __dmul_rn(z, z));
                const double d = __dsqrt_rn(d2);
                if (!(d <= R)) continue;                      // :612
                if (WRITE) {
                    double a2 = __dmul_rn(inv, __dsub_rn(d, R));  // :618
                    if (a2 > 1) a2 = 1;                       // :619
                    fid[o] = gfield[pidx]; fdist[o] = d; facc[o] = a2;
                    o++;
                }
                n++;
            }
        }
    }
    if (!WRITE) counts[i] = n;
}

}  // namespace eut

using namespace eut;

// ---- hooks for cells on a slab-sharded field (slab.cu) ---------------------------------------------------
// A slab's eut_cells holds, of every cell, only the entries that fall on the slab's own points.  What the
// scatter needs per CELL -- the largest distance and sum(max - dist) of the production shares
// (R/run_simulation.R:729-730), the colSums of the uptake shares (:696) -- spans all slabs a cell touches:
// the slab driver combines the partial values and hands the global ones back through these.
namespace eut {

__global__ void __launch_bounds__(256)
k_cell_fsum(const int64_t *__restrict__ cell_ptr, const double *__restrict__ dist, const double *__restrict__ dmax,
            double *__restrict__ fsum, const int n_cells)
{
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= n_cells) return;
    const double m = dmax[w];
    dd s{0.0, 0.0};
    for (int64_t k = cell_ptr[w] + lane; k < cell_ptr[w + 1]; k += 32) dd_add(s, __dsub_rn(m, dist[k]));
    s = dd_warp_sum(s);
    if (lane == 0) fsum[w] = dd_value(s);
}

int64_t cells_count(const eut_cells *c) { return c->M; }

int cells_get_dmax(eut_cells *c, double *host)
{
    EUT_TRY(ensure_device(c->plan->device));
    if (c->M == 0) return EUT_OK;
    EUT_CUDA(cudaStreamSynchronize(0));
    EUT_CUDA(cudaMemcpy(host, c->d_dmax, c->M * sizeof(double), cudaMemcpyDeviceToHost));
    return EUT_OK;
}

// d_dmax <- the global maxima; fsum_host <- this slab's share of sum(max - dist)
int cells_partial_fsum(eut_cells *c, const double *dmax_host, double *fsum_host)
{
    EUT_TRY(ensure_device(c->plan->device));
    if (c->M == 0) return EUT_OK;
    EUT_CUDA(cudaMemcpy(c->d_dmax, dmax_host, c->M * sizeof(double), cudaMemcpyHostToDevice));
    k_cell_fsum<<<(unsigned)((c->M * 32 + 255) / 256), 256, 0, 0>>>(c->d_cell_ptr, c->d_dist, c->d_dmax, c->d_fsum, (int)c->M);
    c->launches++;
    EUT_CUDA(cudaGetLastError());
    EUT_CUDA(cudaMemcpy(fsum_host, c->d_fsum, c->M * sizeof(double), cudaMemcpyDeviceToHost));
    return EUT_OK;
}

int cells_set_fsum(eut_cells *c, const double *fsum_host)
{
    EUT_TRY(ensure_device(c->plan->device));
    if (c->M > 0) EUT_CUDA(cudaMemcpy(c->d_fsum, fsum_host, c->M * sizeof(double), cudaMemcpyHostToDevice));
    c->t_packed = false;                 // the packed (acc.prop, production share) pairs are stale
    return EUT_OK;
}

// the per-cell sums of the last gather (M x C, on the cells' device); nullptr before the first gather
double *cells_fmol(eut_cells *c) { return c->d_fmol; }

// the buffer cells_fmol() points at now holds the sums over ALL slabs, taken from this state of `f`
void cells_mark_gathered(eut_cells *c, const eut_field *f, double field_vol)
{
    c->fmol_field = f; c->fmol_version = f->version; c->fmol_vol = field_vol;
}

bool cells_gather_is_current(const eut_cells *c, const eut_field *f, double field_vol)
{
    return c->d_fmol && c->fmol_field == f && c->fmol_version == f->version && c->fmol_vol == field_vol;
}

}  // namespace eut

extern "C" {

int eut_cells_create(eut_cells **out, eut_plan *plan)
{
    clear_error();
    if (!out || !plan || plan->host_only) { set_error("eut_cells_create: NULL argument or host-only plan"); return EUT_ERR_ARG; }
    *out = nullptr;
    EUT_TRY(ensure_device(plan->device));
    eut_cells *c = new eut_cells();
    c->plan = plan;
    *out = c;
    return EUT_OK;
}

int eut_cells_destroy(eut_cells *c)
{
    if (!c) return EUT_OK;
    cudaSetDevice(c->plan->device);
    cudaDeviceSynchronize();
    cudaFree(c->d_cell_ptr); cudaFree(c->d_fid); cudaFree(c->d_fq); cudaFree(c->d_dist);
    cudaFree(c->d_acc); cudaFree(c->d_ecell); cudaFree(c->d_dmax); cudaFree(c->d_fsum);
    cudaFree(c->d_t_entry); cudaFree(c->d_t_key); cudaFree(c->d_seg_start); cudaFree(c->d_seg_fq);
    cudaFree(c->d_gx); cudaFree(c->d_gy); cudaFree(c->d_gfield); cudaFree(c->d_col_start);
    cudaFree(c->d_col_x); cudaFree(c->d_layer_col); cudaFree(c->d_layer_z); cudaFree(c->d_tmp);
    cudaFree(c->d_fmol); cudaFree(c->d_tfield); cudaFree(c->d_exd); cudaFree(c->d_exmask);
    cudaFree(c->d_pxy); cudaFree(c->d_mom); cudaFree(c->d_cxy); cudaFree(c->d_ex_ptr); cudaFree(c->d_ex_cpd); cudaFree(c->d_ex_flx);
    cudaFree(c->d_ex_cs); cudaFree(c->d_const); cudaFree(c->d_g_idx); cudaFree(c->d_cell_order);
    cudaFree(c->d_t_cell); cudaFree(c->d_t_af); cudaFree(c->d_ex_ptr32); cudaFree(c->d_exrec);
    if (c->ev_ex_ready) cudaEventDestroy(c->ev_ex_ready);
    if (c->ev_ex_free) cudaEventDestroy(c->ev_ex_free);
    delete c;
    return EUT_OK;
}

int eut_cells_set_lists(eut_cells *c, int64_t n_cells, const int64_t *cell_ptr, const int32_t *field_id,
                        const double *field_dist, const double *acc_prop)
{
    clear_error();
    if (!c || n_cells < 0 || !cell_ptr) { set_error("eut_cells_set_lists: bad argument"); return EUT_ERR_ARG; }
    const int64_t T = cell_ptr[n_cells];
    if (cell_ptr[0] != 0 || T < 0 || (T > 0 && (!field_id || !field_dist || !acc_prop)) || T >= INT32_MAX) {
        set_error("eut_cells_set_lists: bad cell_ptr or NULL entry arrays");
        return EUT_ERR_ARG;
    }
    for (int64_t i = 0; i < n_cells; i++)
        if (cell_ptr[i + 1] < cell_ptr[i]) { set_error("eut_cells_set_lists: cell_ptr not monotone"); return EUT_ERR_ARG; }
    EUT_TRY(ensure_device(c->plan->device));
    EUT_TRY(alloc_entries(c, n_cells, T));
    cudaStream_t s = 0;
    EUT_CUDA(cudaMemcpyAsync(c->d_cell_ptr, cell_ptr, (n_cells + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    if (T > 0) {
        EUT_CUDA(cudaMemcpyAsync(c->d_fid, field_id, T * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        EUT_CUDA(cudaMemcpyAsync(c->d_dist, field_dist, T * sizeof(double), cudaMemcpyHostToDevice, s));
        EUT_CUDA(cudaMemcpyAsync(c->d_acc, acc_prop, T * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    return finish_lists(c, s);
}

int eut_cells_build_lists(eut_cells *c, const double *pts, int64_t n_pts, const double *cx,
                          const double *cy, const double *csize, const double *cscv,
                          int64_t n_cells, int dist_mode)
{
    clear_error();
    if (!c || !pts || n_pts != c->plan->N || n_cells < 0 || (n_cells > 0 && (!cx || !cy || !csize || !cscv))) {
        set_error("eut_cells_build_lists: bad argument (pts must be N x 3 for the plan's N)");
        return EUT_ERR_ARG;
    }
    EUT_TRY(ensure_device(c->plan->device));
    cudaStream_t s = 0;
    // (1) geometry index, rebuilt only when the points change (once per simulation)
    const uint64_t h = hash_bytes(pts, (size_t)n_pts * 3 * sizeof(double), 0xabcdull);
    if (h != c->pts_hash || c->n_pts != n_pts || !c->d_gx) {
        std::vector<int32_t> ord(n_pts);
        std::iota(ord.begin(), ord.end(), 0);
        const double *X = pts, *Y = pts + n_pts, *Z = pts + 2 * n_pts;
        std::stable_sort(ord.begin(), ord.end(), [&](int32_t a, int32_t b) {   // setkeyv(z, x, y), :168
            if (Z[a] != Z[b]) return Z[a] < Z[b];
            if (X[a] != X[b]) return X[a] < X[b];
            return Y[a] < Y[b];
        });
        std::vector<double> gx(n_pts), gy(n_pts), col_x, layer_z;
        std::vector<int32_t> gf(n_pts), col_start, layer_col;
        for (int64_t k = 0; k < n_pts; k++) {
            const int32_t i = ord[k];
            gx[k] = X[i]; gy[k] = Y[i]; gf[k] = i + 1;
            const bool new_layer = (k == 0) || Z[i] != Z[ord[k - 1]];
            const bool new_col = new_layer || X[i] != X[ord[k - 1]];
            if (new_layer) { layer_col.push_back((int32_t)col_start.size()); layer_z.push_back(Z[i]); }
            if (new_col) { col_start.push_back((int32_t)k); col_x.push_back(X[i]); }
        }
        layer_col.push_back((int32_t)col_start.size());
        col_start.push_back((int32_t)n_pts);
        cudaFree(c->d_gx); cudaFree(c->d_gy); cudaFree(c->d_gfield); cudaFree(c->d_col_start);
        cudaFree(c->d_col_x); cudaFree(c->d_layer_col); cudaFree(c->d_layer_z);
        c->d_gx = c->d_gy = c->d_col_x = c->d_layer_z = nullptr;
        c->d_gfield = c->d_col_start = c->d_layer_col = nullptr;
        auto up = [&](auto **d, const auto &v) -> int {
            using T = typename std::remove_reference<decltype(v)>::type::value_type;
            EUT_CUDA(cudaMalloc((void **)d, std::max<size_t>(v.size(), 1) * sizeof(T)));
            if (!v.empty()) EUT_CUDA(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
            return EUT_OK;
        };
        EUT_TRY(up(&c->d_gx, gx)); EUT_TRY(up(&c->d_gy, gy)); EUT_TRY(up(&c->d_gfield, gf));
        EUT_TRY(up(&c->d_col_start, col_start)); EUT_TRY(up(&c->d_col_x, col_x));
        EUT_TRY(up(&c->d_layer_col, layer_col)); EUT_TRY(up(&c->d_layer_z, layer_z));
        c->n_cols_g = (int64_t)col_x.size();
        c->n_layers = (int64_t)layer_z.size();
        c->n_pts = n_pts;
        c->pts_hash = h;
    }
    // (2) count, scan, fill
    const int64_t M = n_cells;
    double *d_cell = nullptr;
    EUT_TRY(ensure_tmp(c, (size_t)M * 4 * sizeof(double) + (size_t)(M + 1) * sizeof(int64_t) + 1024));
    d_cell = (double *)c->d_tmp;
    int64_t *d_counts = (int64_t *)(d_cell + 4 * M);
    if (M > 0) {
        EUT_CUDA(cudaMemcpyAsync(d_cell, cx, M * sizeof(double), cudaMemcpyHostToDevice, s));
        EUT_CUDA(cudaMemcpyAsync(d_cell + M, cy, M * sizeof(double), cudaMemcpyHostToDevice, s));
        EUT_CUDA(cudaMemcpyAsync(d_cell + 2 * M, csize, M * sizeof(double), cudaMemcpyHostToDevice, s));
        EUT_CUDA(cudaMemcpyAsync(d_cell + 3 * M, cscv, M * sizeof(double), cudaMemcpyHostToDevice, s));
    }
    std::vector<int64_t> h_ptr(M + 1, 0);
    if (M > 0) {
        const unsigned nb = (unsigned)((M + 127) / 128);
        k_cellmap<false><<<nb, 128, 0, s>>>(c->d_gx, c->d_gy, c->d_gfield, c->d_col_start, c->d_col_x,
                                            c->d_layer_col, c->d_layer_z, (int)c->n_layers, d_cell,
                                            d_cell + M, d_cell + 2 * M, d_cell + 3 * M, (int)M, dist_mode,
                                            d_counts, nullptr, nullptr, nullptr, nullptr);
        EUT_CUDA(cudaGetLastError());
        std::vector<int64_t> h_cnt(M);
        EUT_CUDA(cudaMemcpyAsync(h_cnt.data(), d_counts, M * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        EUT_CUDA(cudaStreamSynchronize(s));
        for (int64_t i = 0; i < M; i++) h_ptr[i + 1] = h_ptr[i] + h_cnt[i];
    }
    const int64_t T = h_ptr[M];
    if (T >= INT32_MAX) { set_error("eut_cells_build_lists: too many entries"); return EUT_ERR_UNSUPPORTED; }
    // the cell arrays live in d_tmp: keep them alive across alloc_entries (which never touches d_tmp)
    EUT_TRY(alloc_entries(c, M, T));
    EUT_CUDA(cudaMemcpyAsync(c->d_cell_ptr, h_ptr.data(), (M + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    if (M > 0 && T > 0) {
        const unsigned nb = (unsigned)((M + 127) / 128);
        k_cellmap<true><<<nb, 128, 0, s>>>(c->d_gx, c->d_gy, c->d_gfield, c->d_col_start, c->d_col_x,
                                           c->d_layer_col, c->d_layer_z, (int)c->n_layers, d_cell,
                                           d_cell + M, d_cell + 2 * M, d_cell + 3 * M, (int)M, dist_mode,
                                           nullptr, c->d_cell_ptr, c->d_fid, c->d_dist, c->d_acc);
        EUT_CUDA(cudaGetLastError());
    }
    c->launches += 2;
    EUT_CUDA(cudaStreamSynchronize(s));
    return finish_lists(c, s);
}

int eut_cells_claim_rescale(eut_cells *c)
{
    clear_error();
    if (!c) return EUT_ERR_ARG;
    EUT_TRY(ensure_device(c->plan->device));
    cudaStream_t s = 0;
    EUT_TRY(build_transposed(c, s));
    c->t_packed = false;                 // acc.prop changes below
    c->fmol_field = nullptr;
    if (c->n_seg > 0) {
        k_claim<<<(unsigned)((c->n_seg + 255) / 256), 256, 0, s>>>(c->d_seg_start, c->d_t_entry, c->d_acc, (int)c->n_seg);
        c->launches++;
        EUT_CUDA(cudaGetLastError());
    }
    EUT_CUDA(cudaStreamSynchronize(s));
    return EUT_OK;
}

int64_t eut_cells_num_entries(const eut_cells *c) { return c ? c->T : 0; }

int eut_cells_get_lists(const eut_cells *c, int64_t *cell_ptr, int32_t *field_id, double *field_dist,
                        double *acc_prop)
{
    clear_error();
    if (!c) return EUT_ERR_ARG;
    EUT_TRY(ensure_device(c->plan->device));
    EUT_CUDA(cudaDeviceSynchronize());
    if (cell_ptr) {
        if (c->d_cell_ptr) EUT_CUDA(cudaMemcpy(cell_ptr, c->d_cell_ptr, (c->M + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost));
        else cell_ptr[0] = 0;
    }
    if (c->T > 0) {
        if (field_id) EUT_CUDA(cudaMemcpy(field_id, c->d_fid, c->T * sizeof(int32_t), cudaMemcpyDeviceToHost));
        if (field_dist) EUT_CUDA(cudaMemcpy(field_dist, c->d_dist, c->T * sizeof(double), cudaMemcpyDeviceToHost));
        if (acc_prop) EUT_CUDA(cudaMemcpy(acc_prop, c->d_acc, c->T * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return EUT_OK;
}

int eut_cells_gather(eut_cells *c, eut_field *f, double field_vol, double *fmol_host)
{
    clear_error();
    if (!c || !f || f->plan != c->plan) { set_error("eut_cells_gather: cells and field must share a plan"); return EUT_ERR_ARG; }
    EUT_TRY(ensure_device(c->plan->device));
    const int64_t n = c->M * f->C;
    if (n == 0) return EUT_OK;
    EUT_TRY(grow(&c->d_fmol, &c->fmol_cap, n));
    EUT_TRY(build_gather_order(c, 0));
    EUT_CUDA(cudaStreamSynchronize(0));  // list construction runs on the legacy stream
    const int64_t n_warps = c->M;
    static const bool old_gather = getenv("EUT_GATHER_V1") != nullptr;
    if (old_gather) {
        constexpr int G = 8, EPL = 5;
        k_gather<G, EPL><<<(unsigned)((n_warps * 32 + 255) / 256), 256, 0, f->stream>>>(
            f->d_buf, c->plan->Np, f->C * c->plan->Np, f->d_cur, c->d_cell_ptr, c->d_g_idx, c->d_fq, c->d_acc,
            field_vol, c->d_fmol, (int)c->M, (int)f->C, c->d_cell_order);
        f->launches++;
    } else {
        const int ldc = (int)((f->C + 31) / 32 * 32);
        EUT_TRY(grow(&c->d_tfield, &c->tfield_cap, c->plan->Np * (int64_t)ldc));
        dim3 tg((unsigned)((c->plan->Np + 31) / 32), (unsigned)(ldc / 32), 1);
        k_field_transpose<<<tg, 256, 0, f->stream>>>(f->d_buf, c->plan->Np, f->C * c->plan->Np, f->d_cur, c->d_tfield,
                                                     (int)c->plan->Np, (int)f->C, ldc);
        k_gather_t<2><<<(unsigned)((n_warps * 32 + 255) / 256), 256, 0, f->stream>>>(
            c->d_tfield, ldc, c->d_cell_ptr, c->d_g_idx, c->d_fq, c->d_acc, field_vol, c->d_fmol, (int)c->M, (int)f->C,
            c->d_cell_order);
        f->launches += 2;
    }
    EUT_CUDA(cudaGetLastError());
    c->fmol_field = f; c->fmol_version = f->version; c->fmol_vol = field_vol;
    if (fmol_host) {
        EUT_CUDA(cudaMemcpyAsync(fmol_host, c->d_fmol, n * sizeof(double), cudaMemcpyDeviceToHost, f->stream));
        EUT_CUDA(cudaStreamSynchronize(f->stream));
    }
    return EUT_OK;
}

int eut_cells_chemotaxis_moments(eut_cells *c, eut_field *f, int32_t col, const double *pts, int64_t n_pts,
                                 const double *cell_x, const double *cell_y, const double *cell_r, double *out)
{
    clear_error();
    if (!c || !f || f->plan != c->plan || !pts || !out || (c->M > 0 && (!cell_x || !cell_y || !cell_r))) {
        set_error("eut_cells_chemotaxis_moments: bad argument");
        return EUT_ERR_ARG;
    }
    if (n_pts != c->plan->N) { set_error("eut_cells_chemotaxis_moments: %lld field points, plan has %lld", (long long)n_pts, (long long)c->plan->N); return EUT_ERR_ARG; }
    if (col < 1 || col > f->C) { set_error("index out of bounds: compound column %d of %lld", col, (long long)f->C); return EUT_ERR_INDEX; }
    EUT_TRY(ensure_device(c->plan->device));
    const int64_t M = c->M, Np = c->plan->Np, N = c->plan->N;
    if (M == 0) return EUT_OK;
    const uint64_t h = hash_bytes(pts, (size_t)N * 2 * sizeof(double), 0xc0ffeeull);
    if (!c->d_pxy || h != c->pxy_hash) {
        std::vector<double> xy((size_t)2 * Np, 0.0);
        for (int64_t i = 0; i < N; i++) { xy[c->plan->perm[i]] = pts[i]; xy[Np + c->plan->perm[i]] = pts[N + i]; }
        cudaFree(c->d_pxy); c->d_pxy = nullptr;
        EUT_CUDA(cudaMalloc((void **)&c->d_pxy, xy.size() * sizeof(double)));
        EUT_CUDA(cudaMemcpy(c->d_pxy, xy.data(), xy.size() * sizeof(double), cudaMemcpyHostToDevice));
        c->pxy_hash = h;
    }
    EUT_TRY(grow(&c->d_mom, &c->mom_cap, 7 * M));
    EUT_TRY(grow(&c->d_cxy, &c->cxy_cap, 3 * M));
    EUT_CUDA(cudaStreamSynchronize(0));            // list construction runs on the legacy stream
    cudaStream_t s = f->stream;
    EUT_CUDA(cudaMemcpyAsync(c->d_cxy, cell_x, M * sizeof(double), cudaMemcpyHostToDevice, s));
    EUT_CUDA(cudaMemcpyAsync(c->d_cxy + M, cell_y, M * sizeof(double), cudaMemcpyHostToDevice, s));
    EUT_CUDA(cudaMemcpyAsync(c->d_cxy + 2 * M, cell_r, M * sizeof(double), cudaMemcpyHostToDevice, s));
    const double *colp = f->d_buf + (int64_t)f->cur[col - 1] * f->C * Np + (int64_t)(col - 1) * Np;
    k_chemo_moments<<<(unsigned)((M * 32 + 255) / 256), 256, 0, s>>>(colp, c->d_pxy, c->d_pxy + Np, c->d_cell_ptr, c->d_fq,
                                                                    c->d_cxy, c->d_mom, (int)M);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    EUT_CUDA(cudaMemcpyAsync(out, c->d_mom, 7 * M * sizeof(double), cudaMemcpyDeviceToHost, s));
    EUT_CUDA(cudaStreamSynchronize(s));
    return EUT_OK;
}

int eut_cells_scatter(eut_cells *c, eut_field *f, double field_vol, const uint8_t *is_constant,
                      const int64_t *ex_ptr, const int32_t *ex_cpd, const double *ex_flx)
{
    clear_error();
    if (!c || !f || f->plan != c->plan || !ex_ptr) { set_error("eut_cells_scatter: bad argument"); return EUT_ERR_ARG; }
    EUT_TRY(ensure_device(c->plan->device));
    const int64_t M = c->M, X = ex_ptr[M];
    if (M == 0 || X == 0 || c->T == 0) return EUT_OK;
    if (!ex_cpd || !ex_flx) { set_error("eut_cells_scatter: NULL exchange arrays"); return EUT_ERR_ARG; }
    {
        std::vector<int64_t> stamp(f->C, -1);   // a cell may list a compound only once (EX_ ids are unique)
        for (int64_t i = 0; i < M; i++)
            for (int64_t e = ex_ptr[i]; e < ex_ptr[i + 1]; e++) {
                if (ex_cpd[e] < 1 || ex_cpd[e] > f->C) {
                    set_error("index out of bounds: exchange %lld refers to compound column %d of %lld", (long long)e + 1, ex_cpd[e], (long long)f->C);
                    return EUT_ERR_INDEX;
                }
                if (stamp[ex_cpd[e] - 1] == i) {
                    set_error("eut_cells_scatter: cell %lld lists compound column %d twice", (long long)i + 1, ex_cpd[e]);
                    return EUT_ERR_ARG;
                }
                stamp[ex_cpd[e] - 1] = i;
            }
    }
    EUT_TRY(build_transposed(c, 0));
    EUT_CUDA(cudaStreamSynchronize(0));
    cudaStream_t s = f->stream;
    EUT_TRY(grow(&c->d_ex_ptr, &c->exptr_cap, M + 1));
    if (X > c->ex_cap || !c->d_ex_cpd) {
        cudaFree(c->d_ex_cpd); cudaFree(c->d_ex_flx); cudaFree(c->d_ex_cs);
        int64_t n = X + X / 8 + 16;
        EUT_CUDA(cudaMalloc((void **)&c->d_ex_cpd, n * sizeof(int32_t)));
        EUT_CUDA(cudaMalloc((void **)&c->d_ex_flx, n * sizeof(double)));
        EUT_CUDA(cudaMalloc((void **)&c->d_ex_cs, n * sizeof(double)));
        c->ex_cap = n;
    }
    if (!c->ev_ex_ready) {
        EUT_CUDA(cudaEventCreateWithFlags(&c->ev_ex_ready, cudaEventDisableTiming));
        EUT_CUDA(cudaEventCreateWithFlags(&c->ev_ex_free, cudaEventDisableTiming));
    }
    cudaStream_t up = f->copy_stream;
    if (c->ex_in_use) EUT_CUDA(cudaStreamWaitEvent(up, c->ev_ex_free, 0));      // the previous scatter has read its lists
    EUT_CUDA(cudaMemcpyAsync(c->d_ex_ptr, ex_ptr, (M + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, up));
    EUT_CUDA(cudaMemcpyAsync(c->d_ex_cpd, ex_cpd, X * sizeof(int32_t), cudaMemcpyHostToDevice, up));
    EUT_CUDA(cudaMemcpyAsync(c->d_ex_flx, ex_flx, X * sizeof(double), cudaMemcpyHostToDevice, up));
    const uint8_t *d_const = nullptr;
    if (is_constant) {
        EUT_TRY(grow(&c->d_const, &c->const_cap, f->C));
        EUT_CUDA(cudaMemcpyAsync(c->d_const, is_constant, f->C, cudaMemcpyHostToDevice, up));
        d_const = c->d_const;
    }
    EUT_CUDA(cudaEventRecord(c->ev_ex_ready, up));
    EUT_CUDA(cudaStreamWaitEvent(s, c->ev_ex_ready, 0));
    struct Release {                                         // whatever path the kernels take below
        eut_cells *c; cudaStream_t s;
        ~Release() { cudaEventRecord(c->ev_ex_free, s); c->ex_in_use = true; }
    } release{c, s};
    const int64_t bs = f->C * c->plan->Np;
    // generation of the scatter kernels (EUT_SCATTER_GEN=1 / 2: the earlier ones, kept as cross-checks)
    int gen = 3;
    if (const char *e = getenv("EUT_SCATTER_GEN")) gen = std::min(3, std::max(1, atoi(e)));
    if (getenv("EUT_SCATTER_V1")) gen = 1;
    if (X >= INT32_MAX) gen = std::min(gen, 2);
    // colSums(fmolPerField) per uptake (R/run_simulation.R:696): when the last gather of these cells ran on
    // exactly this state of the field, its per-cell sums ARE those colSums (the reference takes both from
    // the same fmolPerField matrix, R/scavenge_compounds.R:17-22) -- no second pass over the field
    const bool reuse = gen == 3 && c->fmol_field == f && c->fmol_version == f->version && c->fmol_vol == field_vol &&
                       c->d_fmol && !getenv("EUT_SCATTER_COLSUM");
    if (reuse)
        k_ex_factor<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(c->d_fmol, c->d_ex_ptr, c->d_ex_cpd, c->d_ex_flx, c->d_ex_cs, (int)M,
                                                              field_vol, field_vol / 1e15);
    else
        k_ex_colsum<<<(unsigned)((M * 32 + 255) / 256), 256, 0, s>>>(f->d_buf, c->plan->Np, bs, f->d_cur, c->d_cell_ptr,
                                                                    c->d_fq, c->d_acc, field_vol, c->d_ex_ptr,
                                                                    c->d_ex_cpd, c->d_ex_flx, c->d_ex_cs, (int)M,
                                                                    field_vol / 1e15);
    f->launches++;
    f->version++;
    c->fmol_field = nullptr;
    if (gen == 3) {
        if (!c->t_packed) {
            if (c->T > c->tpack_cap || !c->d_t_cell) {
                cudaFree(c->d_t_cell); cudaFree(c->d_t_af);
                c->d_t_cell = nullptr; c->d_t_af = nullptr;
                const int64_t n = c->T + c->T / 8 + 16;
                EUT_CUDA(cudaMalloc((void **)&c->d_t_cell, n * sizeof(int32_t)));
                EUT_CUDA(cudaMalloc((void **)&c->d_t_af, n * sizeof(double2)));
                c->tpack_cap = n;
            }
            k_t_pack<<<(unsigned)((c->T + 255) / 256), 256, 0, s>>>(c->d_t_entry, c->d_ecell, c->d_acc, c->d_dist, c->d_dmax, c->d_fsum,
                                                                   c->d_t_cell, c->d_t_af, c->T);
            f->launches++;
            c->t_packed = true;
        }
        EUT_TRY(grow(&c->d_ex_ptr32, &c->exptr32_cap, M + 1));
        EUT_TRY(grow(&c->d_exrec, &c->exrec_cap, X));
        k_ex_pack<<<(unsigned)((M + 1 + 255) / 256), 256, 0, s>>>(c->d_ex_ptr, c->d_ex_cpd, c->d_ex_flx, c->d_ex_cs, c->d_ex_ptr32,
                                                                 c->d_exrec, (int)M, (int)f->C);
        f->launches++;
        const unsigned nb = (unsigned)((c->n_seg + 31) / 32);              // 4 warps x 8 points per CTA
        for (int c0 = 0; c0 < f->C; c0 += 64) {
            const int left = (int)f->C - c0;
#define EUT_SCATTER_P(COLS)                                                                                              \
    k_scatter_p<COLS><<<nb, 128, 4 * 2 * (COLS + 1) * 8 * sizeof(double), s>>>(f->d_buf, c->plan->Np, bs, f->d_cur, c->d_seg_start,  \
                                                                     c->d_seg_fq, c->d_t_cell, c->d_t_af, c->d_ex_ptr32,     \
                                                                     c->d_exrec, d_const, (int)c->n_seg, c0, (int)f->C)
            if (left <= 16) EUT_SCATTER_P(16);
            else if (left <= 32) EUT_SCATTER_P(32);
            else EUT_SCATTER_P(64);
#undef EUT_SCATTER_P
            f->launches++;
        }
        EUT_CUDA(cudaGetLastError());
        return EUT_OK;
    }
    if (gen == 2) {
        const int ldc = (int)((f->C + 63) / 64 * 64);
        EUT_TRY(grow(&c->d_exd, &c->exd_cap, M * (int64_t)ldc));
        EUT_TRY(grow(&c->d_exmask, &c->exmask_cap, M * (int64_t)(ldc / 64) * 2));
        EUT_CUDA(cudaMemsetAsync(c->d_exmask, 0, (size_t)M * (ldc / 64) * 2 * sizeof(unsigned long long), s));
        k_ex_dense<<<(unsigned)((M * 32 + 255) / 256), 256, 0, s>>>(c->d_ex_ptr, c->d_ex_cpd, c->d_ex_flx, c->d_ex_cs, c->d_exd,
                                                                   c->d_exmask, (int)M, (int)f->C, ldc);
        k_scatter_t<2><<<(unsigned)((c->n_seg * 32 + 255) / 256), 256, 0, s>>>(
            f->d_buf, c->plan->Np, bs, f->d_cur, c->d_seg_start, c->d_seg_fq, c->d_t_entry, c->d_ecell, c->d_dist,
            c->d_acc, c->d_dmax, c->d_fsum, c->d_exd, c->d_exmask, ldc, d_const, (int)c->n_seg, (int)f->C);
        f->launches += 2;
        EUT_CUDA(cudaGetLastError());
        return EUT_OK;
    }
    constexpr int COLS = 128;             // compound columns per pass (per-warp accumulators: 8 x 128 x 9 B)
    const double vol_l = field_vol / 1e15;
    for (int c0 = 0; c0 < f->C; c0 += COLS) {
        k_scatter<COLS><<<(unsigned)((c->n_seg * 32 + 255) / 256), 256, 0, s>>>(
            f->d_buf, c->plan->Np, bs, f->d_cur, c->d_seg_start, c->d_seg_fq, c->d_t_entry, c->d_ecell,
            c->d_dist, c->d_acc, c->d_dmax, c->d_fsum, field_vol, vol_l, c->d_ex_ptr, c->d_ex_cpd,
            c->d_ex_flx, c->d_ex_cs, d_const, (int)c->n_seg, c0, (int)f->C);
        f->launches++;
    }
    EUT_CUDA(cudaGetLastError());
    // ex_* host arrays may be pageable: the copies above were staged before returning
    return EUT_OK;
}

}  // extern "C"
