/* CPU oracle (TEST INFRASTRUCTURE, never part of the product library): exoenzyme activity and decay,
 * R/run_simulation.R:337-371.
 *
 *   vmax = Kcat * E / 1e6                                                     (:345)
 *   S_t  = Km * lambertW0( S_0/Km * exp((S_0 - vmax * deltaTime * 60 * 60)/Km) )   (:351)
 *   dS   = S_0 - S_t                                                         (:353)
 *   conc[, met_i] += stoich_i * dS   for every non-constant metabolite       (:357-363)
 *   E *= exp(-lambda * deltaTime)                                            (:369-371)
 *
 * lambertW0 comes from the CRAN package lamW (DESCRIPTION:  Imports; not vendored, version not
 * pinned): the principal branch of the Lambert W function to double precision.  It is restated here
 * from its definition -- W e^W = x, W >= -1 -- and solved by Halley's iteration in long double until
 * it stops moving, which is at least as accurate as lamW's Fritsch iteration in double.  PARITY
 * UNPINNED: no golden vectors exist for this step; the parity bar is 1e-12 relative on the resulting
 * concentrations. */
#include <math.h>
#include <stdint.h>

static long double lambert_w0_l(long double x)
{
    if (isnan(x)) return x;
    if (isinf(x)) return x > 0 ? x : NAN;
    const long double inv_e = 0.36787944117144232159552377016146087L;
    if (x < -inv_e) return x > -inv_e - 1e-15L ? -1.0L : NAN;    /* the double nearest to -1/e lies just below it */
    if (x == 0) return x;
    long double w;
    if (x < -0.25L) {                     /* near the branch point */
        const long double p = sqrtl(fmaxl(0.0L, 2 * (2.71828182845904523536028747135266250L * x + 1)));
        w = -1 + p - p * p / 3 + 11 * p * p * p / 72;
    } else if (x < 1) {
        w = x * (1 - x + 1.5L * x * x) / (1 + 0.5L * x * x * x);   /* rough; Halley fixes it */
        if (x > 0.3L) w = log1pl(x) * (1 - log1pl(log1pl(x)) / (2 + log1pl(x)));
    } else {
        const long double l = log1pl(x);
        w = l * (1 - log1pl(l) / (2 + l));
    }
    for (int it = 0; it < 64; it++) {
        const long double ew = expl(w), f = w * ew - x;
        const long double den = ew * (w + 1) - (w + 2) * f / (2 * w + 2);
        if (den == 0 || !isfinite(den)) break;
        const long double wn = w - f / den;
        if (wn == w) break;
        if (fabsl(wn - w) <= 1e-19L * fabsl(wn)) { w = wn; break; }
        w = wn;
    }
    return w;
}

double oracle_lambert_w0(double x) { return (double)lambert_w0_l((long double)x); }

/* conc: n x n_cols column-major (mM), exo: n x n_exo column-major (nM).  mets / stoich: the enzyme's
 * metabolites as 1-based columns of conc (met 1 = the substrate).  In place. */
int oracle_exoenzyme_catalysis(double *conc, int64_t n, int64_t n_cols, const double *exo, int64_t n_exo,
                               int32_t k, double kcat, double km, const int32_t *mets,
                               const double *stoich, int32_t n_mets, const uint8_t *is_const,
                               double delta_time)
{
    if (k < 1 || k > n_exo || n_mets < 1) return 1;
    for (int32_t i = 0; i < n_mets; i++)
        if (mets[i] < 1 || mets[i] > n_cols) return 1;
    const double *e = exo + (int64_t)(k - 1) * n;
    const double *s0col = conc + (int64_t)(mets[0] - 1) * n;
    for (int64_t p = 0; p < n; p++) {
        const double vmax = kcat * e[p] / 1e6;
        const double s0 = s0col[p];
        const double arg = s0 / km * exp((s0 - vmax * delta_time * 60 * 60) / km);
        const double st = km * oracle_lambert_w0(arg);
        const double ds = s0 - st;
        for (int32_t i = 0; i < n_mets; i++)
            if (!is_const || !is_const[mets[i] - 1]) {
                double *c = conc + (int64_t)(mets[i] - 1) * n + p;
                *c = *c + stoich[i] * ds;
            }
    }
    return 0;
}

void oracle_exoenzyme_decay(double *exo, int64_t n, int32_t k, double lambda, double delta_time)
{
    const double f = exp(-lambda * delta_time);
    double *e = exo + (int64_t)(k - 1) * n;
    for (int64_t p = 0; p < n; p++) e[p] = e[p] * f;
}
