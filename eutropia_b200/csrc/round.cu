// Pointwise pieces of a simulation round that sit between the cell exchange and the diffusion
// (SURVEY 8f-1): exoenzyme catalysis (Michaelis-Menten integrated in closed form with the Lambert W
// function, R/run_simulation.R:343-364) and exoenzyme decay (:367-371).  Keeping them on the device
// removes the last reason to download the field every round.
#include <cmath>

#include "common.h"

namespace eut {

namespace {
// Principal branch of the Lambert W function, W e^W = x, for the reference's lamW::lambertW0.
// x >= 0 (the only range the catalysis formula produces): Winitzki-type start, relative error
// < 2e-2, then Fritsch's iteration (4th order): three steps reach binary64 round-off.  -1/e <= x < 0:
// series around the branch point or around 0, then Halley.  Parity bar: 1e-12 relative.
__device__ double lambert_w0(double x)
{
    if (x != x) return x;
    if (x == INFINITY) return x;
    const double inv_e = 0.36787944117144233;
    if (x < -inv_e) return x > -inv_e - 1e-15 ? -1.0 : __longlong_as_double(0x7ff8000000000000ll);
    if (x == 0.0) return x;
    if (x > 0.0) {
        const double l = log1p(x);
        double w = l * (1.0 - log1p(l) / (2.0 + l));
        if (x < 1e-8) return x * (1.0 - x);               // W = x - x^2 + O(x^3)
#pragma unroll
        for (int it = 0; it < 4; it++) {
            const double z = log(x / w) - w;
            const double q = 2.0 * (1.0 + w) * (1.0 + w + 2.0 / 3.0 * z);
            const double e = z / (1.0 + w) * (q - z) / (q - 2.0 * z);
            w = w * (1.0 + e);
        }
        return w;
    }
    double w;
    if (x < -0.25) {
        const double p = sqrt(fmax(0.0, 2.0 * (2.718281828459045 * x + 1.0)));
        w = -1.0 + p - p * p / 3.0 + 11.0 * p * p * p / 72.0;
    } else {
        w = x * (1.0 - x + 1.5 * x * x);
    }
    for (int it = 0; it < 32; it++) {
        const double ew = exp(w), f = w * ew - x;
        const double den = ew * (w + 1.0) - (w + 2.0) * f / (2.0 * w + 2.0);
        if (den == 0.0 || !(den == den)) break;
        const double wn = w - f / den;
        if (wn == w || fabs(wn - w) <= 4e-16 * fabs(wn)) { w = wn; break; }
        w = wn;
    }
    return w;
}

constexpr int kMaxMets = 16;
struct MetArgs {
    long long off[kMaxMets];     // element offset of every updated column inside d_buf
    double stoich[kMaxMets];
    int n;
};

__global__ void __launch_bounds__(256)
k_exo_catalysis(double *__restrict__ cbuf, const double *__restrict__ ecol, const long long s0_off,
                const int32_t *__restrict__ iperm, const int n_padded, const double kcat, const double km,
                const double delta_time, const MetArgs m)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_padded || iperm[q] < 0) return;
    const double vmax = __ddiv_rn(__dmul_rn(kcat, ecol[q]), 1e6);                        // :345
    const double s0 = cbuf[s0_off + q];                                                  // :348
    const double vt = __dmul_rn(__dmul_rn(__dmul_rn(vmax, delta_time), 60.0), 60.0);
    const double arg = __dmul_rn(__ddiv_rn(s0, km), exp(__ddiv_rn(__dsub_rn(s0, vt), km)));
    const double st = __dmul_rn(km, lambert_w0(arg));                                    // :351
    const double ds = __dsub_rn(s0, st);                                                 // :353
    for (int i = 0; i < m.n; i++) {
        double *c = cbuf + m.off[i] + q;
        *c = __dadd_rn(*c, __dmul_rn(m.stoich[i], ds));                                  // :360-361
    }
}

__global__ void __launch_bounds__(256)
k_scale(double *__restrict__ col, const double factor, const int n_padded)
{
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n_padded) col[q] = __dmul_rn(col[q], factor);
}
}  // namespace

int field_exoenzyme_catalysis(eut_field *cf, eut_field *ef, int32_t exo_col, double kcat, double km,
                              const int32_t *met_cols, const double *stoich, int32_t n_mets,
                              const uint8_t *is_constant, double delta_time)
{
    const eut_plan *p = cf->plan;
    if (p->N == 0) return EUT_OK;
    cf->version++;
    MetArgs m;
    m.n = 0;
    const int64_t bs = cf->C * p->Np;
    for (int32_t i = 0; i < n_mets; i++) {
        const int32_t c = met_cols[i] - 1;
        if (is_constant && is_constant[c]) continue;                                     // :359
        m.off[m.n] = (long long)cf->cur[c] * bs + (long long)c * p->Np;
        m.stoich[m.n] = stoich[i];
        m.n++;
    }
    const int32_t c0 = met_cols[0] - 1;
    const long long s0_off = (long long)cf->cur[c0] * bs + (long long)c0 * p->Np;
    const double *ecol = ef->d_buf + (int64_t)ef->cur[exo_col - 1] * ef->C * p->Np + (int64_t)(exo_col - 1) * p->Np;
    // the exoenzyme field may have work in flight on its own stream
    if (ef->stream != cf->stream) EUT_CUDA(cudaStreamSynchronize(ef->stream));
    k_exo_catalysis<<<(unsigned)((p->Np + 255) / 256), 256, 0, cf->stream>>>(cf->d_buf, ecol, s0_off, p->d_iperm,
                                                                           (int)p->Np, kcat, km, delta_time, m);
    cf->launches++;
    EUT_CUDA(cudaGetLastError());
    // ... and whatever the exoenzyme field does next on its own stream (the decay k_scale of
    // R/run_simulation.R:367-371 overwrites the column read here) has to wait for this kernel
    if (ef->stream != cf->stream) {
        cudaEvent_t done = nullptr;
        EUT_CUDA(cudaEventCreateWithFlags(&done, cudaEventDisableTiming));
        EUT_CUDA(cudaEventRecord(done, cf->stream));
        EUT_CUDA(cudaStreamWaitEvent(ef->stream, done, 0));
        EUT_CUDA(cudaEventDestroy(done));          // released by the runtime once the wait has completed
    }
    return EUT_OK;
}

int field_scale_column(eut_field *f, int32_t col0, double factor)
{
    const eut_plan *p = f->plan;
    if (p->N == 0) return EUT_OK;
    f->version++;
    double *dst = f->d_buf + (int64_t)f->cur[col0] * f->C * p->Np + (int64_t)col0 * p->Np;
    k_scale<<<(unsigned)((p->Np + 255) / 256), 256, 0, f->stream>>>(dst, factor, (int)p->Np);
    f->launches++;
    EUT_CUDA(cudaGetLastError());
    return EUT_OK;
}

int max_exo_mets() { return kMaxMets; }

}  // namespace eut
