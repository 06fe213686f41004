// Segment substep kernel: one Jacobi substep of diffChange (src/diffChange.cpp:14-30) with register
// reuse along the raster row.  Tables from segs.cu.
//
//   CTA  = tile x slice of compound columns, warp specialised:
//     * one PRODUCER warp stages, NBUF columns ahead, the tile image of every column (own rows +
//       halo rows, holes left +0.0) with cp.async.bulk (UBLKCP; transaction count on an mbarrier)
//       plus 8-byte cp.async for odd ends / apron cells (cp.async.mbarrier.arrive.noinc on the same
//       mbarrier), and drains the finished columns: ONE cp.async.bulk shared->global per column, so
//       the results cost no LSU wavefronts and are full-sector writes;
//     * CWARPS CONSUMER warps compute.  thread = SPT segments of up to R consecutive points; per
//       segment seven image offsets (the streams O, A..F of segs.cu): 7R+6 LDS.64 per R points
//       instead of 13R.  The R summation chains run in lock step, each strictly in the canonical
//       order A A B B C O- O+ D E E F F (= mat.in order; verified per point by the planner),
//       starting from +0.0.  Streams a whole warp does not have (bottom / top layer) are skipped by
//       a warp-uniform branch; a lane that lacks a stream its warp runs reads the zero header.
//       Results go to a shared-memory result buffer in internal-id order (STS, lane stride R cells).
//   There is no CTA-wide barrier in the column loop: image slots and result buffers are handed
//   over with full/empty mbarriers, so a warp that is ahead (bottom/top-layer warps have 30 % less
//   work) keeps going for up to NBUF-1 columns.
//   generic points (failed the planner's check; rare) use twelve explicit slots read from global.
//
// Arithmetic: only __dadd_rn / __dmul_rn / __dsub_rn, identical to the other substep kernels.
#include <algorithm>

#include "common.h"

namespace eut {

namespace {
__device__ __forceinline__ void s_cp_async_8(uint32_t saddr, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void s_cp_async_16(uint32_t saddr, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(saddr), "l"(g) : "memory");
}
__device__ __forceinline__ void s_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void s_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }
__device__ __forceinline__ void s_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void s_mbar_expect_tx_arrive(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n"
                 "WAIT_%=:\n"
                 " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra DONE_%=;\n"
                 " bra WAIT_%=;\n"
                 "DONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void s_mbar_spin(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n .reg .pred p;\n"
                 "SPIN_%=:\n"
                 " mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 " @p bra SDONE_%=;\n"
                 " bra SPIN_%=;\n"
                 "SDONE_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void s_bulk_g2s(uint32_t saddr, const void *g, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(saddr), "l"(g), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void s_bulk_s2g(void *g, uint32_t saddr, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void s_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void s_fence_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void s_mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void s_cp_async_arrive_noinc(uint32_t bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(bar) : "memory");
}
template <int N>
__device__ __forceinline__ void s_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(N) : "memory"); }
}  // namespace


// debug (dbg & 2048): nanoseconds every CTA of the last launch lived, by (slice, tile); read with eut_debug_tile_ns
__device__ unsigned long long g_tile_ns[2 * 4096];
__device__ __forceinline__ unsigned long long s_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
    return t;
}

template <int R, int SPT, int CWARPS, int NBUF, int OB, int MINB>
__global__ void __launch_bounds__((CWARPS + 1) * 32, MINB)
k_substep_seg(double *__restrict__ buf, const int64_t col_stride, const int64_t buf_stride,
              const int32_t *__restrict__ seg_meta, const int2 *__restrict__ copies,
              const uint4 *__restrict__ segs, const uint4 *__restrict__ gens,
              const double *__restrict__ wout, const int32_t *__restrict__ act_col,
              const int32_t *__restrict__ act_par, const int n_act, const int substep,
              const int img_bytes, const int out_bytes, const int cols_per_cta, const int dbg,
              const int32_t *__restrict__ tile_list)
{
    constexpr int THREADS = (CWARPS + 1) * 32;
    constexpr int CTHREADS = CWARPS * 32;
    constexpr int BE = 4;                                  // bulk copies per producer lane
    constexpr int SE = 16;                                 // single copies per producer lane
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // mbarriers: full[NBUF], empty[NBUF], out_full[OB], out_empty[OB]
    __shared__ __align__(8) unsigned long long s_bar[2 * NBUF + 2 * OB];
    __shared__ long long s_src[kTileMaxCols], s_dst[kTileMaxCols];
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const unsigned long long t_start = (dbg & 2048) ? s_globaltimer() : 0ull;
    asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory");    // the next launch of the stream may start its prologue
    const int tile = tile_list ? __ldg(tile_list + blockIdx.x) : (int)blockIdx.x;
    const int a_begin = blockIdx.y * cols_per_cta;
    const int n_cols = min(n_act, a_begin + cols_per_cta) - a_begin;
    if (n_cols <= 0) return;
    const int4 m0 = __ldg(reinterpret_cast<const int4 *>(seg_meta) + 3 * tile);
    const int4 m1 = __ldg(reinterpret_cast<const int4 *>(seg_meta) + 3 * tile + 1);
    const int4 m2 = __ldg(reinterpret_cast<const int4 *>(seg_meta) + 3 * tile + 2);
    const int t0 = m0.x, cnt = m0.y;
    const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t out_base = s_base + (uint32_t)NBUF * (uint32_t)img_bytes;
    const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(s_bar);
    const uint32_t bar_empty = bar_full + 8u * NBUF;
    const uint32_t bar_ofull = bar_empty + 8u * NBUF;
    const uint32_t bar_oempty = bar_ofull + 8u * OB;

    for (int j = tid; j < n_cols; j += THREADS) {
        const int c = __ldg(act_col + a_begin + j);
        const int b = (__ldg(act_par + a_begin + j) + substep) & 1;
        s_src[j] = (long long)b * buf_stride + (long long)c * col_stride;
        s_dst[j] = (long long)(b ^ 1) * buf_stride + (long long)c * col_stride + t0;
    }
    // the zero header of every staging buffer (absent streams, unused slots, idle lanes read it) stays
    // +0.0 for the whole kernel: the copies never write it.  The result buffers are cleared once as
    // well: their gap cells are never written by a lane and are stored with every column.
    {
        uint4 *z = reinterpret_cast<uint4 *>(smem_raw);
        constexpr int H16 = 8;                              // 128 bytes >= the zero header (R + 3 cells)
        for (int i = tid; i < NBUF * H16; i += THREADS) z[(i / H16) * (img_bytes / 16) + (i % H16)] = make_uint4(0u, 0u, 0u, 0u);
        uint4 *zo = reinterpret_cast<uint4 *>(smem_raw + (size_t)NBUF * img_bytes);
        const int n16 = OB * out_bytes / 16;
        for (int i = tid; i < n16; i += THREADS) zo[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int i = 0; i < NBUF; i++) { s_mbar_init(bar_full + 8u * i, 1 + 32); s_mbar_init(bar_empty + 8u * i, CWARPS); }
        for (int i = 0; i < OB; i++) { s_mbar_init(bar_ofull + 8u * i, CWARPS); s_mbar_init(bar_oempty + 8u * i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    s_fence_async();                                       // generic-proxy zero fill before async-proxy copies
    __syncthreads();
    // Programmatic dependent launch: this CTA may have started while the previous substep's last CTAs
    // were still running (everything above reads only this kernel's own tables); the field itself is
    // touched -- by the producer's copies and stores -- only after the previous grid has completed.
    asm volatile("griddepcontrol.wait;\n" ::: "memory");

    if (tid >= CTHREADS) {
        // ======================= producer warp =================================================
        int bg[BE];
        uint32_t bs[BE];
#pragma unroll
        for (int e = 0; e < BE; e++) {
            bg[e] = 0; bs[e] = 0;
            const int i = lane + e * 32;
            if (i < m0.w) { const int2 d = __ldg(copies + m0.z + i); bg[e] = d.x; bs[e] = (uint32_t)d.y; }
        }
        int sg[SE];
        uint32_t ss[SE];
#pragma unroll
        for (int e = 0; e < SE; e++) {
            const int i = lane + e * 32;
            sg[e] = -1; ss[e] = 0;
            if (i < m1.y) { const int2 d = __ldg(copies + m1.x + i); sg[e] = d.x; ss[e] = (uint32_t)d.y * 8u; }
        }
        __syncwarp();
        const int head = t0 & 1;                           // first point sits at an odd global index
        const int body = (cnt - head) & ~1;
        const int nb_iter = (m0.w + 31) / 32, ns_iter = (m1.y + 31) / 32;
        auto issue = [&](int c) {
            const int slot = c % NBUF;
            if (c >= NBUF) { if (dbg & 256) s_mbar_spin(bar_empty + 8u * slot, (uint32_t)((c / NBUF) - 1) & 1u); else s_mbar_wait(bar_empty + 8u * slot, (uint32_t)((c / NBUF) - 1) & 1u); }
            const double *__restrict__ src = buf + s_src[c];
            const uint32_t sb = s_base + (uint32_t)slot * (uint32_t)img_bytes;
            const uint32_t bar = bar_full + 8u * slot;
            if (lane == 0) s_mbar_expect_tx_arrive(bar, (dbg & 2) ? 0u : (uint32_t)m1.z);
            __syncwarp();
            // the producer shares its scheduler with consumer warps: keep its instruction count down
            // (uniform trip counts instead of BE + SE predicated iterations)
#pragma unroll
            for (int e = 0; e < BE; e++) {
                if (e >= nb_iter) break;
                const uint32_t n = bs[e] >> 16;
                if (n && !(dbg & 2)) s_bulk_g2s(sb + (bs[e] & 0xffffu) * 8u, src + bg[e], n * 8u, bar);
            }
#pragma unroll
            for (int e = 0; e < SE; e++) {
                if (e >= ns_iter) break;
                if (sg[e] >= 0 && !(dbg & 8)) s_cp_async_8(sb + ss[e], src + sg[e]);
            }
            if (dbg & 64) s_mbar_arrive(bar); else s_cp_async_arrive_noinc(bar);
        };
        // Column j is finished when all consumer warps have arrived on out_full[j] (they release the
        // image slot at the same moment).  First refill the freed slot -- the loads are what the
        // consumers will wait for -- then start draining the results; the result buffer of column
        // j-1 is handed back once its store has finished reading shared memory.
        const int pre = min(NBUF, n_cols);
        for (int c = 0; c < pre; c++) issue(c);
        for (int j = 0; j < n_cols; j++) {
            const int ob_i = j % OB;
            { if (dbg & 256) s_mbar_spin(bar_ofull + 8u * ob_i, (uint32_t)(j / OB) & 1u); else s_mbar_wait(bar_ofull + 8u * ob_i, (uint32_t)(j / OB) & 1u); }
            if (j + NBUF < n_cols) issue(j + NBUF);
            double *__restrict__ dst = buf + s_dst[j];
            const uint32_t ob = out_base + (uint32_t)ob_i * (uint32_t)out_bytes;
            const double *res = reinterpret_cast<const double *>(smem_raw + (size_t)NBUF * img_bytes + (size_t)ob_i * out_bytes);
            if (lane == 0) {
                if (body > 0 && !(dbg & 4)) s_bulk_s2g(dst + head, ob + 16u * head, (uint32_t)body * 8u);
                s_bulk_commit();
            }
            if (lane == 1 && head) dst[0] = res[1];
            if (lane == 2 && head + body < cnt) dst[cnt - 1] = res[head + cnt - 1];
            if (j > 0) {
                if (lane == 0 && !(dbg & 128)) s_bulk_wait_read<1>();      // the store of column j-1 has drained its buffer
                __syncwarp();
                if (lane == 0) s_mbar_arrive(bar_oempty + 8u * ((j - 1) % OB));
            }
        }
        if (lane == 0) s_bulk_wait_read<0>();
        if ((dbg & 2048) && lane == 0 && blockIdx.y < 2 && tile < 4096) g_tile_ns[blockIdx.y * 4096 + tile] = s_globaltimer() - t_start;
        return;
    }

    // =========================== consumer warps ================================================
    // segments: byte offsets of the seven streams inside the image, result cell, length, weights
    uint32_t so[SPT][7], sout[SPT], rd[SPT][5];            // rd: packed first | last << 16 of O, A, B, E, F
    int slen[SPT];
    double sw[SPT][R];
    uint32_t lo_mask = 0, up_mask = 0;                     // warp-uniform: some lane has A/B resp. E/F
    const int head = t0 & 1;
#pragma unroll
    for (int s = 0; s < SPT; s++) {
        const int i = tid + s * CTHREADS;
        uint4 r = make_uint4(8u, 0u, 0u, 0u);              // idle lane: O at cell 1, everything else absent
        uint4 r1 = make_uint4(0u, 0u, 0u, 0u), r2 = r1;
        if (i < m2.y) { r = __ldg(segs + 3 * (m2.x + i)); r1 = __ldg(segs + 3 * (m2.x + i) + 1); r2 = __ldg(segs + 3 * (m2.x + i) + 2); }
        so[s][0] = r.x & 0xffffu; so[s][1] = r.x >> 16;
        so[s][2] = r.y & 0xffffu; so[s][3] = r.y >> 16;
        so[s][4] = r.z & 0xffffu; so[s][5] = r.z >> 16;
        so[s][6] = r.w & 0xffffu;
        rd[s][0] = r1.x; rd[s][1] = r1.y; rd[s][2] = r1.z; rd[s][3] = r1.w; rd[s][4] = r2.x;
        sout[s] = ((r.w >> 16) & 0xfffu) * 8u;
        slen[s] = (int)(r.w >> 28);
        const int q0 = t0 + (int)((r.w >> 16) & 0xfffu) - head;
#pragma unroll
        for (int j = 0; j < R; j++) sw[s][j] = (j < slen[s]) ? __ldg(wout + q0 + j) : 0.0;
        lo_mask |= (__any_sync(0xffffffffu, (so[s][1] | so[s][2] | rd[s][1] | rd[s][2]) != 0u) ? 1u : 0u) << s;
        up_mask |= (__any_sync(0xffffffffu, (so[s][5] | so[s][6] | rd[s][3] | rd[s][4]) != 0u) ? 1u : 0u) << s;
    }
    if (dbg & 512) up_mask = 0;                            // ablation: no upper-layer streams (16 of 55 loads less)
    if (dbg & 1024) lo_mask = 0;                           // ablation: no lower-layer streams
    auto s_lds = [](const unsigned char *a) { return *reinterpret_cast<const double *>(a); };
    auto s_sts = [](unsigned char *a, double v) { *reinterpret_cast<double *>(a) = v; };
    for (int j = 0; j < n_cols; j++) {
        const int slot = j % NBUF, ob_i = j % OB;
        { if (dbg & 256) s_mbar_spin(bar_full + 8u * slot, (uint32_t)(j / NBUF) & 1u); else s_mbar_wait(bar_full + 8u * slot, (uint32_t)(j / NBUF) & 1u); }
        const unsigned char *img = smem_raw + (size_t)slot * img_bytes;
        unsigned char *ob = smem_raw + (size_t)NBUF * img_bytes + (size_t)ob_i * out_bytes;
        double res[SPT][R];
#pragma unroll
        for (int s = 0; s < SPT; s++) {
            if (dbg & 1) {
#pragma unroll
                for (int i = 0; i < R; i++) res[s][i] = sw[s][i];
                continue;
            }
            double o[R + 2], acc[R];
            {
                const unsigned char *a = img + so[s][0];
                o[0] = s_lds(img + (rd[s][0] & 0xffffu));
#pragma unroll
                for (int i = 0; i < R; i++) o[i + 1] = s_lds(a + 8u * i);
                o[R + 1] = s_lds(img + (rd[s][0] >> 16));
            }
#pragma unroll
            for (int i = 0; i < R; i++) acc[i] = 0.0;
            auto pair_stream = [&](uint32_t off, uint32_t ends) {
                double v[R + 1];
                const unsigned char *a = img + off;
                v[0] = s_lds(img + (ends & 0xffffu));
#pragma unroll
                for (int i = 1; i < R; i++) v[i] = s_lds(a + 8u * i);
                v[R] = s_lds(img + (ends >> 16));
#pragma unroll
                for (int i = 0; i < R; i++) acc[i] = __dadd_rn(acc[i], v[i]);
#pragma unroll
                for (int i = 0; i < R; i++) acc[i] = __dadd_rn(acc[i], v[i + 1]);
            };
            auto single_stream = [&](uint32_t off) {
                double v[R];
                const unsigned char *a = img + off;
#pragma unroll
                for (int i = 0; i < R; i++) v[i] = s_lds(a + 8u * i);
#pragma unroll
                for (int i = 0; i < R; i++) acc[i] = __dadd_rn(acc[i], v[i]);
            };
            if ((lo_mask >> s) & 1u) { pair_stream(so[s][1], rd[s][1]); pair_stream(so[s][2], rd[s][2]); }
            single_stream(so[s][3]);
#pragma unroll
            for (int i = 0; i < R; i++) acc[i] = __dadd_rn(acc[i], o[i]);          // left neighbour
#pragma unroll
            for (int i = 0; i < R; i++) acc[i] = __dadd_rn(acc[i], o[i + 2]);      // right neighbour
            single_stream(so[s][4]);
            if ((up_mask >> s) & 1u) { pair_stream(so[s][5], rd[s][3]); pair_stream(so[s][6], rd[s][4]); }
#pragma unroll
            for (int i = 0; i < R; i++) {
                double y = __dmul_rn(acc[i], kBDiv);
                y = __dsub_rn(y, __dmul_rn(o[i + 1], sw[s][i]));
                res[s][i] = __dadd_rn(o[i + 1], y);
            }
        }
        if (j >= OB) { if (dbg & 256) s_mbar_spin(bar_oempty + 8u * ob_i, (uint32_t)((j / OB) - 1) & 1u); else s_mbar_wait(bar_oempty + 8u * ob_i, (uint32_t)((j / OB) - 1) & 1u); }
#pragma unroll
        for (int s = 0; s < SPT; s++)
#pragma unroll
            for (int i = 0; i < R; i++)
                if (i < slen[s]) s_sts(ob + sout[s] + 8u * i, res[s][i]);
        // generic points: twelve explicit slots, records re-read per column (L1 resident)
        for (int g = tid; g < m2.w; g += CTHREADS) {
            const uint4 r0 = __ldg(gens + 2 * (m2.z + g)), r1 = __ldg(gens + 2 * (m2.z + g) + 1);
            const uint32_t c[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
            double acc = 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++) {
                acc = __dadd_rn(acc, s_lds(img + (c[k] & 0xffffu) * 8u));
                acc = __dadd_rn(acc, s_lds(img + (c[k] >> 16) * 8u));
            }
            const double cq = s_lds(img + (c[6] & 0xffffu) * 8u);
            const uint32_t oc = c[6] >> 16;
            double y = __dmul_rn(acc, kBDiv);
            y = __dsub_rn(y, __dmul_rn(cq, __ldg(wout + t0 + (int)oc - head)));
            s_sts(ob + oc * 8u, __dadd_rn(cq, y));
        }
        if (!(dbg & 32)) s_fence_async();        // results (generic proxy) before the producer's bulk store (async proxy)
        __syncwarp();
        if (lane == 0) {
            s_mbar_arrive(bar_empty + 8u * slot);          // this warp is done reading the image slot
            s_mbar_arrive(bar_ofull + 8u * ob_i);          // ... and its results of column j are in the buffer
        }
    }
}

template <int R, int SPT, int CWARPS, int NBUF, int OB, int MINB>
static int launch_seg_t(eut_field *f, int n_act, int substep)
{
    const eut_plan *p = f->plan;
    const SegPlanHost &S = p->segs;
    if (S.R != R || S.max_tile_segs > SPT * CWARPS * 32 || S.max_bulk > 128 || S.max_single > 512) {
        set_error("segment kernel: plan has R=%d, %d segments per tile; kernel shape is R=%d, %d", S.R,
                  S.max_tile_segs, R, SPT * CWARPS * 32);
        return EUT_ERR_UNSUPPORTED;
    }
    const int img_bytes = ((S.img_points * 8 + 127) / 128) * 128;
    const int out_bytes = (((S.max_tile_points + 2) * 8 + 127) / 128) * 128;
    const size_t smem = (size_t)NBUF * img_bytes + (size_t)OB * out_bytes;
    if (smem > 225 * 1024) {
        set_error("segment kernel: %zu bytes of shared memory per CTA", smem);
        return EUT_ERR_UNSUPPORTED;
    }
    // slab overlap (halo.cu): only the boundary or only the interior tiles
    const int32_t *tile_list = f->opt_phase == 1 ? f->d_tiles_b : (f->opt_phase == 2 ? f->d_tiles_i : nullptr);
    const int n_tiles = f->opt_phase == 1 ? f->n_tiles_b : (f->opt_phase == 2 ? f->n_tiles_i : p->n_tiles);
    if (n_tiles == 0) return EUT_OK;
    // slices of columns per tile: ~6 waves of CTAs keep the tail of the last wave small, but a CTA
    // costs ~4 us before its first column (tables, first image), so it should get at least 16 columns
    // Launches that run side by side with another lane's (field.cu) have their last wave filled by the
    // other lane: fewer, longer CTAs win there (32 columns per CTA: 0.235 ms per substep against
    // 0.245 ms with 16, 1 M points x 64 columns).
    int slices = 1;
    const int64_t want = (f->in_lanes ? 2ll : 6ll) * 148 * MINB;
    const int min_cols = f->in_lanes ? 32 : 16;
    while ((int64_t)n_tiles * slices < want && slices < n_act && (n_act + slices) / (slices + 1) >= min_cols) slices++;
    // Small fields (configs[1]: 73 tiles): not even one wave of CTAs, so a launch lasts as long as its
    // longest CTA and most SMs idle -- split the columns further, down to 4 per CTA (100 k points x 16
    // columns: 19.4 -> 10.5 us per substep, profiles/r02_square_sweep.txt)
    while ((int64_t)n_tiles * slices < 148ll * MINB && slices < n_act && (n_act + slices) / (slices + 1) >= 4) slices++;
    if (f->opt_cpt > 0) slices = (n_act + f->opt_cpt - 1) / f->opt_cpt;
    slices = std::max(slices, (n_act + kTileMaxCols - 1) / kTileMaxCols);
    const int cols_per_cta = (n_act + slices - 1) / slices;
    dim3 grid((unsigned)n_tiles, (unsigned)((n_act + cols_per_cta - 1) / cols_per_cta), 1);
    auto kern = k_substep_seg<R, SPT, CWARPS, NBUF, OB, MINB>;
    static bool attr_set[64] = {false};
    if (!attr_set[p->device & 63]) {
        EUT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 225 * 1024));
        attr_set[p->device & 63] = true;
    }
    // programmatic dependent launch (EUT_PDL=0 turns it off): the next substep's CTAs are placed as soon as
    // SMs fall free and run their prologue (tables, zero fill, mbarrier init) under this launch's tail;
    // they wait in griddepcontrol.wait until this grid has completed before they touch the field
    static const bool pdl = [] { const char *e = getenv("EUT_PDL"); return !e || atoi(e) != 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3((CWARPS + 1) * 32, 1, 1); cfg.dynamicSmemBytes = smem; cfg.stream = f->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    const int64_t col_stride = p->Np, buf_stride = f->C * p->Np;
    EUT_CUDA(cudaLaunchKernelEx(&cfg, kern, f->d_buf, col_stride, buf_stride, (const int32_t *)p->d_seg_meta, (const int2 *)p->d_copies,
                                (const uint4 *)p->d_segs, (const uint4 *)p->d_gens, (const double *)p->d_w, f->act_col, f->act_par,
                                n_act, substep, img_bytes, out_bytes, cols_per_cta, f->opt_debug, tile_list));
    return EUT_OK;
}

}  // namespace eut
extern "C" EUT_API int eut_debug_tile_ns(unsigned long long *out, int n)
{
    return cudaMemcpyFromSymbol(out, eut::g_tile_ns, sizeof(unsigned long long) * (size_t)std::min(n, 2 * 4096)) == cudaSuccess ? 0 : 1;
}
namespace eut {

// Kernel shapes: (R, segments per thread, consumer warps, image slots, result buffers, CTAs per SM).
// The plan's tile capacity (EUT_SEG_TILE, default 224 segments) picks the shape.
int launch_seg(eut_field *f, int n_act, int substep)
{
    const SegPlanHost &S = f->plan->segs;
    const int nb = f->opt_nbuf == 2 ? 2 : (f->opt_nbuf == 4 ? 4 : 3);
    const int shape = f->opt_shape;
    if (S.R == 7) {
        if (S.max_tile_segs <= 224) {
            if (shape == 1) return launch_seg_t<7, 1, 7, 3, 2, 2>(f, n_act, substep);
            if (shape == 5) return launch_seg_t<7, 1, 7, 6, 3, 1>(f, n_act, substep);
            if (shape == 6) return launch_seg_t<7, 1, 7, 4, 4, 1>(f, n_act, substep);
            if (nb == 2) return launch_seg_t<7, 1, 7, 2, 2, 2>(f, n_act, substep);
            if (nb == 4) return launch_seg_t<7, 1, 7, 4, 2, 2>(f, n_act, substep);
            return launch_seg_t<7, 1, 7, 3, 3, 2>(f, n_act, substep);
        }
        if (S.max_tile_segs <= 448) {      // large fields: one CTA per SM, 14 consumer warps (halo / own 0.18)
            if (shape == 1) return launch_seg_t<7, 1, 14, 3, 2, 1>(f, n_act, substep);
            if (shape == 2) return launch_seg_t<7, 1, 14, 4, 2, 1>(f, n_act, substep);
            if (shape == 3) return launch_seg_t<7, 1, 14, 4, 3, 1>(f, n_act, substep);
            if (shape == 5) return launch_seg_t<7, 1, 14, 3, 4, 1>(f, n_act, substep);
            return launch_seg_t<7, 1, 14, 3, 3, 1>(f, n_act, substep);
        }
    }
    if (S.R == 5 && S.max_tile_segs <= 224) return launch_seg_t<5, 1, 7, 3, 2, 2>(f, n_act, substep);
    if (S.R == 9 && S.max_tile_segs <= 224) return launch_seg_t<9, 1, 7, 3, 2, 2>(f, n_act, substep);
    set_error("segment kernel: no compiled shape for R=%d with %d segments per tile", S.R, S.max_tile_segs);
    return EUT_ERR_UNSUPPORTED;
}

}  // namespace eut
