// Device kernels of the diffusion substep (src/diffChange.cpp:16-29 == :50-63 per column) and the
// layout conversion between the caller's column-major reference order and the internal order.
//
// Arithmetic contract (bit-exact with the reference on finite inputs):
//   acc = +0.0;  for each in-neighbour in mat.in row order: acc = acc + c[from]     (DADD.RN)
//   y   = acc * (1/13)                                                               (DMUL.RN)
//   y   = y - c[p] * ((nn_p - 1) * (1/13))                                           (DMUL, DADD)
//   c'[p] = c[p] + y                                                                 (DADD.RN)
// Only __dadd_rn/__dmul_rn/__dsub_rn are used so that nvcc can never contract to FMA.
// The reference zeroes its scratch vector with `ychg = ychg * 0`; on finite inputs the result of
// the substep is the same as starting from +0.0 ("signed zeros"), which is what the
// Jacobi gather form does here.  Adding +0.0 for an unused neighbour slot is exact as well: the
// running sum starts at +0.0 and x + (+0.0) == x for every x except -0.0, which a sum that started
// at +0.0 can never be.
#include <algorithm>

#include "common.h"

namespace eut {

// ---------------------------------------------------------------------------------------------
// ELL path: thread = one point, CPT compound columns per thread (neighbour indices are loaded
// once and reused for all CPT columns); warp = 32 consecutive internal positions.  Used when the
// graph does not tile (tiles.cu) and as a cross-check of the tiled kernel.
// ---------------------------------------------------------------------------------------------
template <int W, int CPT>
__global__ void __launch_bounds__(256)
k_substep_ell(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
              const int32_t *__restrict__ nbr, const uint8_t *__restrict__ deg,
              const double *__restrict__ wout, const int32_t *__restrict__ act_col,
              const int32_t *__restrict__ act_par, const int n_act, const int substep,
              const int n_points)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_points) return;
    const int a0 = blockIdx.y * CPT;
    const int d = deg[q];
    int idx[W];
#pragma unroll
    for (int k = 0; k < W; k++) idx[k] = (k < d) ? __ldg(nbr + (int64_t)k * col_stride + q) : q;
    const double wq = __ldg(wout + q);

#pragma unroll
    for (int j = 0; j < CPT; j++) {
        const int a = a0 + j;
        if (a >= n_act) break;
        const int c = __ldg(act_col + a);
        const int b = (__ldg(act_par + a) + substep) & 1;
        const double *__restrict__ src = buf + (int64_t)b * buf_stride + (int64_t)c * col_stride;
        double *__restrict__ dst = buf + (int64_t)(b ^ 1) * buf_stride + (int64_t)c * col_stride;
        const double cq = __ldg(src + q);
        double v[W];
#pragma unroll
        for (int k = 0; k < W; k++) v[k] = (k < d) ? __ldg(src + idx[k]) : 0.0;
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < W; k++)
            if (k < d) acc = __dadd_rn(acc, v[k]);
        double y = __dmul_rn(acc, kBDiv);
        y = __dsub_rn(y, __dmul_rn(cq, wq));
        dst[q] = __dadd_rn(cq, y);
    }
}

// ---------------------------------------------------------------------------------------------
// Generic path: any in-degree (CSR) and any number of mat.out rows per point, in row order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_substep_csr(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
              const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
              const int64_t *__restrict__ out_ptr, const double *__restrict__ out_w,
              const int32_t *__restrict__ act_col, const int32_t *__restrict__ act_par,
              const int n_act, const int substep, const int n_points)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int a = blockIdx.y;
    if (q >= n_points || a >= n_act) return;
    const int c = act_col[a];
    const int b = (act_par[a] + substep) & 1;
    const double *__restrict__ src = buf + (int64_t)b * buf_stride + (int64_t)c * col_stride;
    double *__restrict__ dst = buf + (int64_t)(b ^ 1) * buf_stride + (int64_t)c * col_stride;
    const double cq = src[q];
    double acc = 0.0;
    for (int64_t k = row_ptr[q]; k < row_ptr[q + 1]; k++) acc = __dadd_rn(acc, src[col[k]]);
    double y = __dmul_rn(acc, kBDiv);
    for (int64_t k = out_ptr[q]; k < out_ptr[q + 1]; k++) y = __dsub_rn(y, __dmul_rn(cq, out_w[k]));
    dst[q] = __dadd_rn(cq, y);
}

template <int W>
static int launch_ell_w(eut_field *f, int n_act, int substep)
{
    const eut_plan *p = f->plan;
    int cpt = f->opt_cpt > 0 ? f->opt_cpt : (n_act >= 8 ? 8 : (n_act >= 4 ? 4 : (n_act >= 2 ? 2 : 1)));
    const int block = 256;
    // all Np internal positions: a plan may leave gaps between its points (segs.cu pads ragged row
    // pieces); gap cells have degree 0 and weight 0 and keep their value
    dim3 grid((unsigned)((p->Np + block - 1) / block), 1, 1);
    const int64_t bs = f->C * p->Np;
    const int32_t *act_col = f->act_col, *act_par = f->act_par;
#define EUT_LAUNCH_ELL(CPT)                                                                      \
    grid.y = (unsigned)((n_act + CPT - 1) / CPT);                                                \
    k_substep_ell<W, CPT><<<grid, block, 0, f->stream>>>(f->d_buf, p->Np, bs, p->d_nbr, p->d_deg, \
                                                         p->d_w, act_col, act_par, n_act, substep, \
                                                         (int)p->Np)
    if (cpt >= 8) { EUT_LAUNCH_ELL(8); }
    else if (cpt >= 4) { EUT_LAUNCH_ELL(4); }
    else if (cpt >= 2) { EUT_LAUNCH_ELL(2); }
    else { EUT_LAUNCH_ELL(1); }
#undef EUT_LAUNCH_ELL
    return EUT_OK;
}

int launch_substep(eut_field *f, int n_act, int substep)
{
    const eut_plan *p = f->plan;
    if (n_act <= 0 || p->N == 0) return EUT_OK;
    const int64_t bs = f->C * p->Np;
    if (p->segmented && f->opt_variant != 1) {
        EUT_TRY(launch_seg(f, n_act, substep));                           // stencil_seg.cu
    } else if (p->tiled && f->opt_variant != 1) {
        EUT_TRY(launch_tile_u(f, n_act, substep));                        // stencil_tile.cu
    } else if (p->ell_w > 0) {
        if (p->ell_w <= 4) EUT_TRY(launch_ell_w<4>(f, n_act, substep));
        else if (p->ell_w <= 8) EUT_TRY(launch_ell_w<8>(f, n_act, substep));
        else EUT_TRY(launch_ell_w<12>(f, n_act, substep));
    } else {
        const int block = 256;
        dim3 grid((unsigned)((p->N + block - 1) / block), (unsigned)n_act, 1);
        k_substep_csr<<<grid, block, 0, f->stream>>>(f->d_buf, p->Np, bs, p->d_row_ptr, p->d_col,
                                                     p->d_out_ptr, p->d_out_w, f->act_col,
                                                     f->act_par, n_act, substep,
                                                     (int)p->N);
    }
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

// ---------------------------------------------------------------------------------------------
// Layout conversion.  Stage buffers hold whole columns in reference order (what R hands over);
// the field holds them in internal order with stride Np.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_permute_in(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
             const double *__restrict__ stage, const int64_t ld, const int32_t *__restrict__ iperm,
             const int32_t *__restrict__ cols, const int32_t *__restrict__ cur, const int n_padded)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_padded) return;
    const int j = blockIdx.y;
    const int c = cols[j];
    const int i = iperm[q];
    buf[(int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + q] =
        (i >= 0) ? stage[(int64_t)j * ld + i] : 0.0;
}

// round(x, digits) as base R (>= 4.0.0) does it for digits > 0 (nmath/fround.c, restated from its published
// description -- base R is not part of the reference tree): of the two decimal neighbours
// floor(|x| * 10^d) / 10^d and ceil(|x| * 10^d) / 10^d the one closer to x AS REPRESENTED, ties to the even
// integer; values with more than 15 significant digits before the cut (and NaN / Inf) are returned as they
// are.  R compares the two distances in long double; here they are compared through an exact quotient
// (double-double, Dekker products without FMA): round(0.15, 1) = 0.1, round(0.45, 1) = 0.5 as ?Round
// documents.  The host mirror (recording.r_round) runs the same IEEE operations and gives the same bits.
// Used by the concentration recording, R/run_simulation.R:514.
__device__ __forceinline__ void r_two_prod(const double a, const double b, double &p, double &e)
{
    p = __dmul_rn(a, b);
    const double c = 134217729.0;                           // 2^27 + 1
    double t = __dmul_rn(c, a);
    const double ah = __dsub_rn(t, __dsub_rn(t, a)), al = __dsub_rn(a, ah);
    t = __dmul_rn(c, b);
    const double bh = __dsub_rn(t, __dsub_rn(t, b)), bl = __dsub_rn(b, bh);
    e = __dadd_rn(__dadd_rn(__dadd_rn(__dsub_rn(__dmul_rn(ah, bh), p), __dmul_rn(ah, bl)), __dmul_rn(al, bh)), __dmul_rn(al, bl));
}
__device__ __forceinline__ void r_exact_quotient(const double n, const double d, double &hi, double &lo)
{
    hi = __ddiv_rn(n, d);
    double p, e;
    r_two_prod(hi, d, p, e);
    lo = __ddiv_rn(__dsub_rn(__dsub_rn(n, p), e), d);
}
__device__ __forceinline__ double r_round_digits(const double x, const double p10)
{
    const double ax = fabs(x);
    if (!(ax < 1e15 / p10)) return x;
    const double x10 = __dmul_rn(p10, ax), i10 = floor(x10), c10 = ceil(x10);
    double xd, xd_lo, xu, xu_lo;
    r_exact_quotient(i10, p10, xd, xd_lo);
    r_exact_quotient(c10, p10, xu, xu_lo);
    const double dd = __dsub_rn(__dsub_rn(ax, xd), xd_lo), du = __dadd_rn(__dsub_rn(xu, ax), xu_lo);
    const bool even = fmod(i10, 2.0) == 0.0;
    return copysign((dd < du || (dd == du && even)) ? xd : xu, x);
}

__global__ void __launch_bounds__(256)
k_permute_out(const double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
              double *__restrict__ stage, const int64_t ld, const int32_t *__restrict__ perm,
              const int32_t *__restrict__ cols, const int32_t *__restrict__ cur, const int n_points,
              const double round_p10)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const int j = blockIdx.y;
    const int c = cols[j];
    double v = buf[(int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + perm[i]];
    if (round_p10 > 0.0) v = r_round_digits(v, round_p10);
    stage[(int64_t)j * ld + i] = v;
}

int launch_permute_in(eut_field *f, const double *d_stage, int64_t ld, const int32_t *d_cols,
                      const int32_t *d_cur, int64_t n_cols, cudaStream_t s)
{
    const eut_plan *p = f->plan;
    if (n_cols <= 0) return EUT_OK;
    dim3 grid((unsigned)((p->Np + 255) / 256), (unsigned)n_cols, 1);
    k_permute_in<<<grid, 256, 0, s>>>(f->d_buf, p->Np, f->C * p->Np, d_stage, ld, p->d_iperm, d_cols,
                                      d_cur, (int)p->Np);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

int launch_permute_out(eut_field *f, double *d_stage, int64_t ld, const int32_t *d_cols,
                       const int32_t *d_cur, int64_t n_cols, cudaStream_t s)
{
    const eut_plan *p = f->plan;
    if (n_cols <= 0 || p->N == 0) return EUT_OK;
    dim3 grid((unsigned)((p->N + 255) / 256), (unsigned)n_cols, 1);
    k_permute_out<<<grid, 256, 0, s>>>(f->d_buf, p->Np, f->C * p->Np, d_stage, ld, p->d_perm, d_cols,
                                       d_cur, (int)p->N, f->round_p10);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

// d_cur[col[i]] = value[i]: the device mirror of "which buffer holds column c" after a diffuse call
__global__ void k_set_cur(int32_t *__restrict__ cur, const int32_t *__restrict__ col,
                          const int32_t *__restrict__ value, const int n)
{
    for (int i = threadIdx.x; i < n; i += blockDim.x) cur[col[i]] = value[i];
}

int launch_set_cur(eut_field *f, const int32_t *d_col, const int32_t *d_value, int n)
{
    if (n <= 0) return EUT_OK;
    k_set_cur<<<1, 128, 0, f->stream>>>(f->d_cur, d_col, d_value, n);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

// ---------------------------------------------------------------------------------------------
// Column reductions (min, max, sum) and fill: diffuse_compoundsPar's `sd > 0` column selection and
// its D = Inf -> column-mean rule (R/growthEnvironment.R:372, :393-401).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_column_stats(const double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
               const int32_t *__restrict__ cur, const int32_t *__restrict__ iperm, const int n_padded,
               const int n_points, double *__restrict__ stats, const int n_cols)
{
    // one block per column; deterministic tree reduction (fixed block size, fixed stride)
    const int c = blockIdx.x;
    const double *src = buf + (int64_t)cur[c] * buf_stride + (int64_t)c * col_stride;
    double mn = INFINITY, mx = -INFINITY, sm = 0.0;
    bool nan = false;
    for (int q = threadIdx.x; q < n_padded; q += blockDim.x) {
        if (iperm[q] < 0) continue;                        // padding / gap cell
        double v = src[q];
        nan |= (v != v);
        mn = fmin(mn, v); mx = fmax(mx, v); sm += v;
    }
    __shared__ double s_mn[256], s_mx[256], s_sm[256];
    __shared__ int s_nan[256];
    s_mn[threadIdx.x] = mn; s_mx[threadIdx.x] = mx; s_sm[threadIdx.x] = sm; s_nan[threadIdx.x] = nan;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            s_mn[threadIdx.x] = fmin(s_mn[threadIdx.x], s_mn[threadIdx.x + o]);
            s_mx[threadIdx.x] = fmax(s_mx[threadIdx.x], s_mx[threadIdx.x + o]);
            s_sm[threadIdx.x] += s_sm[threadIdx.x + o];
            s_nan[threadIdx.x] |= s_nan[threadIdx.x + o];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        stats[c] = s_nan[0] ? qnan : s_mn[0];
        stats[n_cols + c] = s_nan[0] ? qnan : s_mx[0];
        stats[2 * n_cols + c] = n_points > 0 ? s_sm[0] / (double)n_points : qnan;
    }
}

__global__ void __launch_bounds__(256)
k_fill(double *__restrict__ dst, const double value, const int n_points)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_points) dst[q] = value;
}

int launch_column_stats(eut_field *f, const int32_t *d_cur)
{
    const eut_plan *p = f->plan;
    if (f->C == 0) return EUT_OK;
    k_column_stats<<<(unsigned)f->C, 256, 0, f->stream>>>(f->d_buf, p->Np, f->C * p->Np, d_cur, p->d_iperm,
                                                          (int)p->Np, (int)p->N, f->d_stats, (int)f->C);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

int launch_fill_column(eut_field *f, int32_t col0, double value)
{
    const eut_plan *p = f->plan;
    if (p->N == 0) return EUT_OK;
    f->version++;
    double *dst = f->d_buf + (int64_t)f->cur[col0] * f->C * p->Np + (int64_t)col0 * p->Np;
    k_fill<<<(unsigned)((p->Np + 255) / 256), 256, 0, f->stream>>>(dst, value, (int)p->Np);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

}  // namespace eut
