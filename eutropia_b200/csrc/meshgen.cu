// Neighbour index maps on the device: `build.DCM` + the emit step of `initialize`
// (R/growthEnvironment.R:164-273, :138-149) as sort / search joins.
//
//   coordinates -> sorted unique values per axis -> integer cell key per point -> points sorted by key;
//   for every point and each of the 12 rhombic-dodecahedral offsets (:184-197) the shifted coordinate
//   is snapped per axis like `correct_numbers` does (:199-232: rolling join "next observation carried
//   backward" onto the unique values, i.e. the smallest value >= q, rejected when it is farther than
//   field.size / 10 away), looked up in the sorted keys (:239-250), and emitted as the directed edge
//   (from = the point, to = the owner of the snapped coordinate) -- DCM[id, neighbour] = 1/13,
//   transposed, `which(DCM != 0)` (:262-265, :140);  edges are sorted by (to, from) and made unique
//   (:141-142);  n_neighbors = out-degree + 1 for the diagonal (:143).
//
// Integer output, so the parity bar is bit-exact equality with the host mirror (mesh.build_dcm),
// which the CPU test-suite pins against hand-built meshes.  Floating-point inputs are only shifted
// and compared; --fmad=false keeps `x + d * fs` the two IEEE operations the reference performs.
#include <cub/cub.cuh>

#include "common.h"

namespace eut {

namespace {
__constant__ double c_off[12][3] = {
    {0.0, 1.0, 0.0}, {1.0, 0.0, 0.0}, {0.0, -1.0, 0.0}, {-1.0, 0.0, 0.0},
    {-0.5, 0.5, 1.0}, {0.5, 0.5, 1.0}, {0.5, -0.5, 1.0}, {-0.5, -0.5, 1.0},
    {-0.5, 0.5, -1.0}, {0.5, 0.5, -1.0}, {0.5, -0.5, -1.0}, {-0.5, -0.5, -1.0}};

__device__ __forceinline__ int lower_bound_d(const double *a, int n, double q)
{
    int lo = 0, hi = n;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (a[mid] < q) lo = mid + 1; else hi = mid; }
    return lo;
}
__device__ __forceinline__ int64_t lower_bound_k(const int64_t *a, int64_t n, int64_t q)
{
    int64_t lo = 0, hi = n;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (a[mid] < q) lo = mid + 1; else hi = mid; }
    return lo;
}
// correct_numbers: index of the smallest unique value >= q, or -1 (none, or farther than tol)
__device__ __forceinline__ int snap(const double *u, int n, double q, double tol)
{
    const int i = lower_bound_d(u, n, q);
    if (i >= n) return -1;
    if (fabs(__dsub_rn(u[i], q)) > tol) return -1;
    return i;
}

__global__ void __launch_bounds__(256)
k_point_keys(const double *x, const double *y, const double *z, const double *ux, int nux, const double *uy, int nuy,
             const double *uz, int nuz, int64_t *key, int32_t *id, int n)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int ix = lower_bound_d(ux, nux, x[p]), iy = lower_bound_d(uy, nuy, y[p]), iz = lower_bound_d(uz, nuz, z[p]);
    key[p] = ((int64_t)iz * nuy + iy) * nux + ix;
    id[p] = p;
}

__global__ void __launch_bounds__(256)
k_edges(const double *x, const double *y, const double *z, const double *ux, int nux, const double *uy, int nuy,
        const double *uz, int nuz, const int64_t *skey, const int32_t *order, const double fs, const double tol,
        int64_t *pair, int n)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * 12) return;
    const int p = (int)(t / 12), o = (int)(t % 12);
    int64_t out = INT64_MAX;
    const int sx = snap(ux, nux, __dadd_rn(x[p], __dmul_rn(c_off[o][0], fs)), tol);
    const int sy = snap(uy, nuy, __dadd_rn(y[p], __dmul_rn(c_off[o][1], fs)), tol);
    const int sz = snap(uz, nuz, __dadd_rn(z[p], __dmul_rn(c_off[o][2], fs)), tol);
    if (sx >= 0 && sy >= 0 && sz >= 0) {
        const int64_t k = ((int64_t)sz * nuy + sy) * nux + sx;
        const int64_t pos = lower_bound_k(skey, n, k);
        if (pos < n && skey[pos] == k) out = (int64_t)(order[pos] + 1) * ((int64_t)n + 1) + (p + 1);   // to * (n+1) + from
    }
    pair[t] = out;
}

__global__ void __launch_bounds__(256)
k_decode(const int64_t *pair, int64_t n_edges, int64_t n1, int32_t *from, int32_t *to, int32_t *outdeg)
{
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_edges) return;
    const int64_t v = pair[e];
    const int32_t t = (int32_t)(v / n1), f = (int32_t)(v % n1);
    from[e] = f; to[e] = t;
    atomicAdd(outdeg + (f - 1), 1);
}

__global__ void __launch_bounds__(256)
k_mat_out(const int32_t *outdeg, int32_t *id, int32_t *cnt, int n)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    id[p] = p + 1; cnt[p] = outdeg[p] + 1;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    template <typename T> T *as() { return static_cast<T *>(p); }
};

// sorted unique values of one coordinate column; returns the count through *n_out
int unique_axis(const double *d_col, int n, DevBuf &out, int *n_out, DevBuf &tmp, size_t &tmp_bytes)
{
    DevBuf sorted, cnt;
    EUT_CUDA(sorted.alloc((size_t)n * 8));
    EUT_CUDA(out.alloc((size_t)n * 8));
    EUT_CUDA(cnt.alloc(8));
    size_t need = 0, need2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, need, d_col, sorted.as<double>(), n);
    cub::DeviceSelect::Unique(nullptr, need2, sorted.as<double>(), out.as<double>(), cnt.as<int>(), n);
    need = std::max(need, need2);
    if (need > tmp_bytes) { if (tmp.p) cudaFree(tmp.p); tmp.p = nullptr; EUT_CUDA(tmp.alloc(need)); tmp_bytes = need; }
    EUT_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, need, d_col, sorted.as<double>(), n));
    EUT_CUDA(cub::DeviceSelect::Unique(tmp.p, need, sorted.as<double>(), out.as<double>(), cnt.as<int>(), n));
    EUT_CUDA(cudaMemcpy(n_out, cnt.p, sizeof(int), cudaMemcpyDeviceToHost));
    return EUT_OK;
}
}  // namespace

int mesh_build_dcm(const double *pts, int64_t ld, int64_t n64, double fs, int device, int32_t *mat_in, int64_t cap,
                   int64_t *n_edges_out, int32_t *mat_out)
{
    EUT_TRY(ensure_device(device));
    if (n64 >= (int64_t)INT32_MAX / 12) { set_error("eut_build_dcm: %lld points are too many", (long long)n64); return EUT_ERR_UNSUPPORTED; }
    const int n = (int)n64;
    *n_edges_out = 0;
    if (n == 0) return EUT_OK;
    DevBuf dx, dy, dz, ux, uy, uz, key, id, skey, order, pair, spair, tmp, cnt, dfrom, dto, deg, oid, ocnt;
    size_t tmp_bytes = 0;
    EUT_CUDA(dx.alloc((size_t)n * 8)); EUT_CUDA(dy.alloc((size_t)n * 8)); EUT_CUDA(dz.alloc((size_t)n * 8));
    EUT_CUDA(cudaMemcpy(dx.p, pts, (size_t)n * 8, cudaMemcpyHostToDevice));
    EUT_CUDA(cudaMemcpy(dy.p, pts + ld, (size_t)n * 8, cudaMemcpyHostToDevice));
    EUT_CUDA(cudaMemcpy(dz.p, pts + 2 * ld, (size_t)n * 8, cudaMemcpyHostToDevice));
    int nux = 0, nuy = 0, nuz = 0;
    EUT_TRY(unique_axis(dx.as<double>(), n, ux, &nux, tmp, tmp_bytes));
    EUT_TRY(unique_axis(dy.as<double>(), n, uy, &nuy, tmp, tmp_bytes));
    EUT_TRY(unique_axis(dz.as<double>(), n, uz, &nuz, tmp, tmp_bytes));
    const unsigned nb = (unsigned)((n + 255) / 256);
    EUT_CUDA(key.alloc((size_t)n * 8)); EUT_CUDA(skey.alloc((size_t)n * 8));
    EUT_CUDA(id.alloc((size_t)n * 4)); EUT_CUDA(order.alloc((size_t)n * 4));
    k_point_keys<<<nb, 256>>>(dx.as<double>(), dy.as<double>(), dz.as<double>(), ux.as<double>(), nux, uy.as<double>(), nuy,
                              uz.as<double>(), nuz, key.as<int64_t>(), id.as<int32_t>(), n);
    {
        size_t need = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, need, key.as<int64_t>(), skey.as<int64_t>(), id.as<int32_t>(), order.as<int32_t>(), n);
        if (need > tmp_bytes) { if (tmp.p) cudaFree(tmp.p); tmp.p = nullptr; EUT_CUDA(tmp.alloc(need)); tmp_bytes = need; }
        EUT_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, need, key.as<int64_t>(), skey.as<int64_t>(), id.as<int32_t>(), order.as<int32_t>(), n));
    }
    const int64_t n12 = (int64_t)n * 12;
    EUT_CUDA(pair.alloc((size_t)n12 * 8)); EUT_CUDA(spair.alloc((size_t)n12 * 8));
    k_edges<<<(unsigned)((n12 + 255) / 256), 256>>>(dx.as<double>(), dy.as<double>(), dz.as<double>(), ux.as<double>(), nux,
                                                   uy.as<double>(), nuy, uz.as<double>(), nuz, skey.as<int64_t>(),
                                                   order.as<int32_t>(), fs, fs / 10, pair.as<int64_t>(), n);
    EUT_CUDA(cnt.alloc(8));
    {
        size_t need = 0, need2 = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, need, pair.as<int64_t>(), spair.as<int64_t>(), (int)n12);
        cub::DeviceSelect::Unique(nullptr, need2, spair.as<int64_t>(), pair.as<int64_t>(), cnt.as<int>(), (int)n12);
        need = std::max(need, need2);
        if (need > tmp_bytes) { if (tmp.p) cudaFree(tmp.p); tmp.p = nullptr; EUT_CUDA(tmp.alloc(need)); tmp_bytes = need; }
        EUT_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, need, pair.as<int64_t>(), spair.as<int64_t>(), (int)n12));
        EUT_CUDA(cub::DeviceSelect::Unique(tmp.p, need, spair.as<int64_t>(), pair.as<int64_t>(), cnt.as<int>(), (int)n12));
    }
    int n_unique = 0;
    EUT_CUDA(cudaMemcpy(&n_unique, cnt.p, sizeof(int), cudaMemcpyDeviceToHost));
    // the "no neighbour" sentinel sorts last and survives Unique once
    int64_t last = 0;
    if (n_unique > 0) EUT_CUDA(cudaMemcpy(&last, pair.as<int64_t>() + (n_unique - 1), 8, cudaMemcpyDeviceToHost));
    const int64_t E = n_unique - ((n_unique > 0 && last == INT64_MAX) ? 1 : 0);
    *n_edges_out = E;
    if (E > cap) { set_error("eut_build_dcm: %lld edges do not fit the %lld rows provided", (long long)E, (long long)cap); return EUT_ERR_ARG; }
    EUT_CUDA(dfrom.alloc((size_t)E * 4)); EUT_CUDA(dto.alloc((size_t)E * 4)); EUT_CUDA(deg.alloc((size_t)n * 4));
    EUT_CUDA(cudaMemset(deg.p, 0, (size_t)n * 4));
    if (E > 0)
        k_decode<<<(unsigned)((E + 255) / 256), 256>>>(pair.as<int64_t>(), E, (int64_t)n + 1, dfrom.as<int32_t>(), dto.as<int32_t>(), deg.as<int32_t>());
    EUT_CUDA(oid.alloc((size_t)n * 4)); EUT_CUDA(ocnt.alloc((size_t)n * 4));
    k_mat_out<<<nb, 256>>>(deg.as<int32_t>(), oid.as<int32_t>(), ocnt.as<int32_t>(), n);
    EUT_CUDA(cudaGetLastError());
    if (E > 0) {
        EUT_CUDA(cudaMemcpy(mat_in, dfrom.p, (size_t)E * 4, cudaMemcpyDeviceToHost));
        EUT_CUDA(cudaMemcpy(mat_in + cap, dto.p, (size_t)E * 4, cudaMemcpyDeviceToHost));
    }
    EUT_CUDA(cudaMemcpy(mat_out, oid.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    EUT_CUDA(cudaMemcpy(mat_out + n, ocnt.p, (size_t)n * 4, cudaMemcpyDeviceToHost));
    return EUT_OK;
}

}  // namespace eut
