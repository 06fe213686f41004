// Tiled substep kernel (tables from tiles.cu): the fallback of the segment kernel (stencil_seg.cu) for
// plans the segment planner declines, and EUT_ORDER_TILED.  One CTA = one compact tile of <= 256*PPT
// points and a slice of the active compound columns.  For every column the tile's image -- own points
// plus halo runs -- is staged into shared memory NBUF columns deep: 16-byte aligned runs by
// cp.async.bulk (UBLKCP, issued by warp 0, completing on an mbarrier with a transaction count), odd
// ends and lone halo points by 8-byte cp.async; one __syncthreads per column.  Per thread:
//
//   * the image offsets of the 12 neighbour slots are kept UNPACKED, one 32-bit register each, so a
//     neighbour load is a single `LDS.64 R, [Roff + URimage]` (no unpack, no address add);
//   * the degree class (<= 8 / > 8 in-neighbours: bottom+top layers vs middle layers of the
//     rhombic-dodecahedral mesh) is a warp-uniform branch decided once per tile (vote at setup),
//     not a per-load predicate; unused slots of a point read the zero cell, which is exact;
//   * the staging slot is a rotating counter (no modulo), all column bookkeeping is uniform.
//
// 13 shared-memory loads per update: bound by the LSU data pipe (profiles/r01_v6_tile_lean_ncu.csv),
// which is what the segment kernel's register reuse addresses.
#include <algorithm>

#include "common.h"

namespace eut {

namespace {
__device__ __forceinline__ void t_cp_async_8(uint32_t saddr, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void t_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void t_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void t_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void t_mbar_expect_tx_arrive(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void t_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n"
                 "WAIT_%=:\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra DONE_%=;\n"
                 " bra WAIT_%=;\n"
                 "DONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void t_bulk_g2s(uint32_t saddr, const void *g, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(saddr), "l"(g), "r"(bytes), "r"(bar) : "memory");
}
}  // namespace

template <int PPT, int THREADS, int NBUF, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
k_substep_tile_u(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
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

    // copy descriptors: bulk copies live in the lanes of warp 0 (two each), single 8-byte copies in
    // the first threads of the CTA (SE each)
    int bg[2] = {0, 0};
    uint32_t bs[2] = {0, 0};
    if (tid < 32) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int i = tid + e * 32;
            if (i < m0.w) { const int2 d = __ldg(copies + m0.z + i); bg[e] = d.x; bs[e] = (uint32_t)d.y; }
        }
    }
    constexpr int SE = (512 + THREADS - 1) / THREADS;
    int sg[SE];
    uint32_t ss[SE];
#pragma unroll
    for (int e = 0; e < SE; e++) {
        const int i = tid + e * THREADS;
        sg[e] = -1; ss[e] = 0;
        if (i < m1.y) { const int2 d = __ldg(copies + m1.x + i); sg[e] = d.x; ss[e] = (uint32_t)d.y * 8u; }
    }
    // per point: 12 unpacked image offsets (bytes), own offset, outflow weight, degree class
    uint32_t off[PPT][W], own[PPT];
    double wq[PPT];
    uint32_t act_mask = 0, hi_any = 0;
#pragma unroll
    for (int p = 0; p < PPT; p++) {
        const int lp = tid + p * THREADS;
        const bool on = lp < cnt;
        const int64_t q = t0 + (on ? lp : 0);
        const int d = on ? (int)__ldg(deg + q) : 0;
        act_mask |= (on ? 1u : 0u) << p;
        hi_any |= (__any_sync(0xffffffffu, d > 8) ? 1u : 0u) << p;     // warp-uniform
        own[p] = on ? (uint32_t)(2 + (t0 & 1) + lp) * 8u : 0u;
        wq[p] = on ? __ldg(wout + q) : 0.0;
#pragma unroll
        for (int k = 0; k < W; k++) off[p][k] = on ? (uint32_t)__ldg(loc16 + (int64_t)k * col_stride + q) : 0u;
    }
    for (int j = tid; j < n_cols; j += THREADS) {
        const int c = __ldg(act_col + a_begin + j);
        const int b = (__ldg(act_par + a_begin + j) + substep) & 1;
        s_src[j] = (long long)b * buf_stride + (long long)c * col_stride;
        s_dst[j] = (long long)(b ^ 1) * buf_stride + (long long)c * col_stride + t0;
    }
    if (tid < NBUF) {
        double *z = reinterpret_cast<double *>(smem_raw + (size_t)tid * img_bytes);
        z[0] = 0.0; z[1] = 0.0;                // the zero cells: target of unused neighbour slots
        t_mbar_init(bar_base + 8u * tid, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();

    auto issue = [&](int j, int slot) {
        if (j < n_cols) {
            const double *__restrict__ src = buf + s_src[j];
            const uint32_t sb = s_base + (uint32_t)slot * (uint32_t)img_bytes;
            if (tid < 32) {
                const uint32_t bar = bar_base + 8u * slot;
                if (tid == 0) t_mbar_expect_tx_arrive(bar, (uint32_t)m1.z);
                __syncwarp();
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const uint32_t n = bs[e] >> 16;
                    if (n) t_bulk_g2s(sb + (bs[e] & 0xffffu) * 8u, src + bg[e], n * 8u, bar);
                }
            }
#pragma unroll
            for (int e = 0; e < SE; e++)
                if (sg[e] >= 0) t_cp_async_8(sb + ss[e], src + sg[e]);
        }
        t_cp_async_commit();
    };

    int slot_issue = 0;
#pragma unroll
    for (int s = 0; s < NBUF - 1; s++) { issue(s, slot_issue); slot_issue = (slot_issue + 1 == NBUF) ? 0 : slot_issue + 1; }
    int slot = 0;
    uint32_t phase = 0;
    for (int j = 0; j < n_cols; j++) {
        t_cp_async_wait<NBUF - 2>();
        t_mbar_wait(bar_base + 8u * slot, phase);
        __syncthreads();
        issue(j + NBUF - 1, slot_issue);
        slot_issue = (slot_issue + 1 == NBUF) ? 0 : slot_issue + 1;
        double *__restrict__ dst = buf + s_dst[j];
        const unsigned char *img = smem_raw + (size_t)slot * img_bytes;
#pragma unroll
        for (int p = 0; p < PPT; p++) {
            double v[W];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = *reinterpret_cast<const double *>(img + off[p][k]);
            const double cq = *reinterpret_cast<const double *>(img + own[p]);
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 8; k++) acc = __dadd_rn(acc, v[k]);
            if ((hi_any >> p) & 1u) {
#pragma unroll
                for (int k = 8; k < W; k++) v[k] = *reinterpret_cast<const double *>(img + off[p][k]);
#pragma unroll
                for (int k = 8; k < W; k++) acc = __dadd_rn(acc, v[k]);
            }
            double y = __dmul_rn(acc, kBDiv);
            y = __dsub_rn(y, __dmul_rn(cq, wq[p]));
            if ((act_mask >> p) & 1u) dst[tid + p * THREADS] = __dadd_rn(cq, y);
        }
        slot++;
        if (slot == NBUF) { slot = 0; phase ^= 1u; }
    }
}

template <int PPT, int THREADS, int NBUF, int MINB>
static int launch_u(eut_field *f, int n_act, int substep)
{
    const eut_plan *p = f->plan;
    const int img_bytes = ((p->tiles.img_points * 8 + 127) / 128) * 128;
    const size_t smem = (size_t)NBUF * img_bytes;
    int slices = 1;
    const int64_t want = 4ll * 148 * MINB;
    while ((int64_t)p->n_tiles * slices < want && slices < n_act) slices++;
    if (f->opt_cpt > 0) slices = (n_act + f->opt_cpt - 1) / f->opt_cpt;
    slices = std::max(slices, (n_act + kTileMaxCols - 1) / kTileMaxCols);
    const int cols_per_cta = (n_act + slices - 1) / slices;
    dim3 grid((unsigned)p->n_tiles, (unsigned)((n_act + cols_per_cta - 1) / cols_per_cta), 1);
    auto kern = k_substep_tile_u<PPT, THREADS, NBUF, MINB>;
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

// Thread shapes: the plan's largest tile decides how many points a thread must take.
int launch_tile_u(eut_field *f, int n_act, int substep)
{
    const int mtp = f->plan->tiles.max_tile_points;
    const int shape = f->opt_shape;
    if (shape == 1 && mtp <= 4 * 384) return launch_u<4, 384, 3, 1>(f, n_act, substep);
    if (shape == 2 && mtp <= 3 * 512) return launch_u<3, 512, 3, 1>(f, n_act, substep);
    if (shape == 3 && mtp <= 5 * 256) return launch_u<5, 256, 3, 2>(f, n_act, substep);
    if (shape == 4 && mtp <= 3 * 256) return launch_u<3, 256, 3, 3>(f, n_act, substep);
    if (shape == 5 && mtp <= 2 * 256) return launch_u<2, 256, 3, 4>(f, n_act, substep);
    if (shape == 6 && mtp <= 4 * 256) return launch_u<4, 256, 2, 2>(f, n_act, substep);
    if (mtp <= 4 * 256) return launch_u<4, 256, 3, 2>(f, n_act, substep);
    if (mtp <= 4 * 384) return launch_u<4, 384, 3, 1>(f, n_act, substep);
    set_error("tiled kernel: tile of %d points exceeds the compiled thread shapes", mtp);
    return EUT_ERR_UNSUPPORTED;
}

}  // namespace eut
