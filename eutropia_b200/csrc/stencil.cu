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

// ---------------------------------------------------------------------------------------------
// Tiled path (tiles.cu builds the tables): one CTA = one compact tile of <= 256*PPT points and a
// slice of the active compound columns.  For every column the tile's image -- own points plus halo
// runs -- is staged into shared memory NBUF columns deep:
//   * the 16-byte aligned runs by cp.async.bulk (TMA engine, SASS UBLKCP), issued by the lanes of
//     warp 0 and completing on an mbarrier with a transaction count;
//   * odd run ends and lone halo points by 8-byte cp.async (LDGSTS) from the threads that own them.
// HBM/L2 latency is hidden by the copies of the next columns, not by occupancy.  The per-thread
// state (copy descriptors, 16-bit image offsets of the 12 neighbour slots of PPT points, outflow
// weights) is loaded once per tile and reused for all columns of the slice.  The compute step
// issues all 13 shared-memory loads of two points before their two DADD chains start, so that the
// chains (fixed summation order, hence serial) overlap each other and the loads.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_8_if(bool on, uint32_t saddr, const void *g)
{
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %0, 0;\n"
                 " @p cp.async.ca.shared.global [%1], [%2], 8;\n}\n" ::"r"((int)on), "r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_arrive(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n"
                 "WAIT_%=:\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra DONE_%=;\n"
                 " bra WAIT_%=;\n"
                 "DONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t saddr, const void *g, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(saddr), "l"(g), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ double lds_f64(uint32_t saddr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ double lds_f64_if(bool on, uint32_t saddr)
{
    double v;
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %1, 0;\n mov.f64 %0, 0d0000000000000000;\n"
                 " @p ld.shared.f64 %0, [%2];\n}\n" : "=d"(v) : "r"((int)on), "r"(saddr));
    return v;
}

template <int PPT, int THREADS, int NBUF, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_substep_tile(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
               const int32_t *__restrict__ tile_meta, const int2 *__restrict__ copies,
               const uint16_t *__restrict__ loc16, const uint8_t *__restrict__ deg,
               const double *__restrict__ wout, const int32_t *__restrict__ act_col,
               const int32_t *__restrict__ act_par, const int n_act, const int substep,
               const int img_bytes, const int cols_per_cta)
{
    constexpr int W = 12;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long s_bar[NBUF];
    __shared__ long long s_src[kTileMaxCols], s_dst[kTileMaxCols];
    const int tid = threadIdx.x;
    const int tile = blockIdx.x;
    const int a_begin = blockIdx.y * cols_per_cta;
    const int n_cols = min(n_act, a_begin + cols_per_cta) - a_begin;
    if (n_cols <= 0) return;
    const int4 m0 = __ldg(reinterpret_cast<const int4 *>(tile_meta) + 2 * tile);
    const int4 m1 = __ldg(reinterpret_cast<const int4 *>(tile_meta) + 2 * tile + 1);
    const int t0 = m0.x, cnt = m0.y;
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t bar_base = (uint32_t)__cvta_generic_to_shared(s_bar);

    // bulk copies: lanes of warp 0 hold up to two each; single copies: every thread up to two
    int bg[2] = {0, 0};
    uint32_t bs[2] = {0, 0};
    if (tid < 32) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int i = tid + e * 32;
            if (i < m0.w) { const int2 d = __ldg(copies + m0.z + i); bg[e] = d.x; bs[e] = (uint32_t)d.y; }
        }
    }
    constexpr int SE = (512 + THREADS - 1) / THREADS;   // single copies per thread (<= 512 per tile)
    int sg[SE];
    uint32_t ss[SE];
    bool son[SE];
#pragma unroll
    for (int e = 0; e < SE; e++) {
        sg[e] = 0; ss[e] = 0;
        const int i = tid + e * THREADS;
        son[e] = i < m1.y;
        if (son[e]) { const int2 d = __ldg(copies + m1.x + i); sg[e] = d.x; ss[e] = (uint32_t)d.y * 8u; }
    }
    // neighbour offsets (bytes inside the image), two per register; degree class; outflow weight
    uint32_t lo[PPT][W / 2];
    double wq[PPT];
    uint32_t own_q[PPT];
    uint32_t act_mask = 0, hi_mask = 0;
#pragma unroll
    for (int p = 0; p < PPT; p++) {
        const int lp = tid + p * THREADS;
        const bool on = lp < cnt;
        const int64_t q = t0 + (on ? lp : 0);
        const int d = (int)__ldg(deg + q);
        act_mask |= (on ? 1u : 0u) << p;
        hi_mask |= ((on && d > 8) ? 1u : 0u) << p;
        own_q[p] = on ? (uint32_t)(2 + (t0 & 1) + lp) * 8u : 0u;   // own value inside the image
        wq[p] = __ldg(wout + q);
#pragma unroll
        for (int k = 0; k < W / 2; k++) {
            const uint32_t a = __ldg(loc16 + (int64_t)(2 * k) * col_stride + q);
            const uint32_t b = __ldg(loc16 + (int64_t)(2 * k + 1) * col_stride + q);
            lo[p][k] = on ? (a | (b << 16)) : 0u;   // idle lanes read the zero cell
        }
    }
    // per-column source / destination offsets of this slice (doubles from `buf`)
    for (int j = tid; j < n_cols; j += THREADS) {
        const int c = __ldg(act_col + a_begin + j);
        const int b = (__ldg(act_par + a_begin + j) + substep) & 1;
        s_src[j] = (long long)b * buf_stride + (long long)c * col_stride;
        s_dst[j] = (long long)(b ^ 1) * buf_stride + (long long)c * col_stride;
    }
    if (tid < NBUF) {
        // image cells 0,1 of every buffer stay zero: the target of unused neighbour slots
        double *z = reinterpret_cast<double *>(smem_raw + (size_t)tid * img_bytes);
        z[0] = 0.0; z[1] = 0.0;
        mbar_init(bar_base + 8u * tid, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int j) {
        if (j < n_cols) {
            const double *__restrict__ src = buf + s_src[j];
            const int slot = j % NBUF;
            const uint32_t sb = s_base + (uint32_t)slot * (uint32_t)img_bytes;
            const uint32_t bar = bar_base + 8u * slot;
            if (tid < 32) {
                if (tid == 0) mbar_expect_tx_arrive(bar, (uint32_t)m1.z);
                __syncwarp();
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t n = bs[e] >> 16;
                    if (n) bulk_g2s(sb + (bs[e] & 0xffffu) * 8u, src + bg[e], n * 8u, bar);
                }
            }
#pragma unroll
            for (int e = 0; e < SE; e++)   // warps without single copies skip the block
                if ((tid & ~31) + e * THREADS < m1.y) cp_async_8_if(son[e], sb + ss[e], src + sg[e]);
        }
        cp_async_commit();
    };

#pragma unroll
    for (int s = 0; s < NBUF - 1; s++) issue(s);
    for (int j = 0; j < n_cols; j++) {
        // column j has landed: own 8-byte copies (all groups but the newest NBUF-2), bulk copies
        // (mbarrier phase), then the block barrier makes every thread's copies visible to all.
        // The same barrier proves that everybody is done reading buffer (j-1) % NBUF, which is the
        // one column j+NBUF-1 is staged into next -- one barrier per column.
        cp_async_wait<NBUF - 2>();
        const int slot = j % NBUF;
        mbar_wait(bar_base + 8u * slot, (uint32_t)((j / NBUF) & 1));
        __syncthreads();
        issue(j + NBUF - 1);
        double *__restrict__ dst = buf + s_dst[j] + t0;
        // Per point: all shared-memory loads, then the summation chain in mat.in order (serial by
        // contract); ptxas interleaves the chains of the PPT points as far as registers allow.
        // plain shared-memory loads from a uniform image base: ptxas folds base + offset into the
        // [R + UR] addressing form, one LDS per neighbour after a one-instruction unpack
        const unsigned char *img = smem_raw + (size_t)slot * img_bytes;
#pragma unroll
        for (int p = 0; p < PPT; p++) {
            const bool on = (act_mask >> p) & 1u, hi = (hi_mask >> p) & 1u;
            double v[W];
#pragma unroll
            for (int k = 0; k < W; k++) {
                const uint32_t off = (k & 1) ? (lo[p][k >> 1] >> 16) : (lo[p][k >> 1] & 0xffffu);
                if (k < 8 || hi) v[k] = *reinterpret_cast<const double *>(img + off);
            }
            const double cq = *reinterpret_cast<const double *>(img + own_q[p]);
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < W; k++)
                if (k < 8 || hi) acc = __dadd_rn(acc, v[k]);
            double y = __dmul_rn(acc, kBDiv);
            y = __dsub_rn(y, __dmul_rn(cq, wq[p]));
            if (on) dst[tid + p * THREADS] = __dadd_rn(cq, y);
        }
    }
}

template <int PPT, int THREADS, int NBUF, int MINB>
static int launch_tile_t(eut_field *f, int n_act, int substep)
{
    const eut_plan *p = f->plan;
    const int img_bytes = ((p->tiles.img_points * 8 + 127) / 128) * 128;
    const size_t smem = (size_t)NBUF * img_bytes;
    // enough CTAs to fill the machine a few times over, as few column slices as possible (the
    // per-tile tables are re-read once per slice)
    int slices = 1;
    const int64_t want = 4ll * 148 * 2;
    while ((int64_t)p->n_tiles * slices < want && slices < n_act) slices++;
    if (f->opt_cpt > 0) slices = (n_act + f->opt_cpt - 1) / f->opt_cpt;
    slices = std::max(slices, (n_act + kTileMaxCols - 1) / kTileMaxCols);
    const int cols_per_cta = (n_act + slices - 1) / slices;
    dim3 grid((unsigned)p->n_tiles, (unsigned)((n_act + cols_per_cta - 1) / cols_per_cta), 1);
    auto kern = k_substep_tile<PPT, THREADS, NBUF, MINB>;
    static bool attr_set[64] = {false};
    if (!attr_set[p->device & 63]) {
        EUT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[p->device & 63] = true;
    }
    kern<<<grid, THREADS, smem, f->stream>>>(f->d_buf, p->Np, f->C * p->Np, p->d_tile_meta, p->d_copies,
                                             p->d_loc16, p->d_deg, p->d_w, f->act_col, f->act_par,
                                             n_act, substep, img_bytes, cols_per_cta);
    return EUT_OK;
}

// Thread shape of the tiled kernel: tiles hold up to 1536 points = THREADS x PPT.  Default: 512
// threads x 3 points, 64 registers, 2 CTAs (32 warps) per SM; option "ppt" selects the others.
int launch_tile(eut_field *f, int n_act, int substep)
{
    const int mtp = f->plan->tiles.max_tile_points;
    switch (f->opt_ppt) {
    case 6: return launch_tile_t<6, 256, 3, 2>(f, n_act, substep);
    case 4: if (mtp <= 1024) return launch_tile_t<4, 256, 3, 3>(f, n_act, substep);
            return launch_tile_t<4, 384, 3, 2>(f, n_act, substep);
    case 2: return launch_tile_t<2, 768, 3, 1>(f, n_act, substep);
    default: return launch_tile_t<3, 512, 3, 2>(f, n_act, substep);
    }
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
        if (f->opt_ppt > 0) EUT_TRY(launch_tile(f, n_act, substep));      // first-generation kernel
        else EUT_TRY(launch_tile_u(f, n_act, substep));                   // stencil_tile.cu
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

__global__ void __launch_bounds__(256)
k_permute_out(const double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
              double *__restrict__ stage, const int64_t ld, const int32_t *__restrict__ perm,
              const int32_t *__restrict__ cols, const int32_t *__restrict__ cur, const int n_points)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_points) return;
    const int j = blockIdx.y;
    const int c = cols[j];
    stage[(int64_t)j * ld + i] = buf[(int64_t)cur[c] * buf_stride + (int64_t)c * col_stride + perm[i]];
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
                                       d_cur, (int)p->N);
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
    double *dst = f->d_buf + (int64_t)f->cur[col0] * f->C * p->Np + (int64_t)col0 * p->Np;
    k_fill<<<(unsigned)((p->Np + 255) / 256), 256, 0, f->stream>>>(dst, value, (int)p->Np);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

}  // namespace eut
