/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by or called from the product
 * library (eutropia_b200/csrc).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it, and only as the checker / the timed CPU arm.
 *
 * PARITY PIN: the reference ships no tests, golden vectors or fixtures for this path (SURVEY.md
 * section 4 and 8c), but its one compiled source file builds: oracle/Makefile target `ref` compiles
 * /root/reference/src/diffChange.cpp, unmodified and from where it lies, against a stand-in for the
 * few Armadillo types it uses (oracle/ref_shim/RcppArmadillo.h; R, Rcpp and Armadillo themselves are
 * not installed) into oracle/_ref/libeutropia_ref.so.  tests/test_oracle.py holds this restatement
 * to that library bit for bit (finite, signed-zero, Inf/NaN data, shuffled edge lists, bounds
 * errors) and to the vectors it wrote (tests/golden/diffusion_reference.npz), and the GPU tests
 * compare the CUDA path with both.  What the pin cannot cover is Armadillo itself: the stand-in
 * restates the documented element-wise meaning of `v * s`, `M.col(j) += v` and the checked
 * operator(); see its header.
 *
 * What is restated (all citations relative to /root/reference):
 *   eut_oracle_diffchange      <- src/diffChange.cpp:6-35   (diffChange: double indices, serial)
 *   eut_oracle_diffchange_par  <- src/diffChange.cpp:38-69  (diffChangePar: int indices, OpenMP)
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp (see oracle/Makefile).  -ffp-contract=off because R's
 * default x86-64 flags carry no FMA target, so the reference never fuses a*b+c.
 *
 * Conventions shared with the reference: matrices are column-major, indices are 1-based.
 * Return value: 0 ok; >0 = the condition under which Armadillo's bounds check (enabled: no
 * ARMA_NO_DEBUG in src/Makevars) would raise and Rcpp would turn it into an R error.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define EUT_ORACLE_OK 0
#define EUT_ORACLE_ERR_INDEX 1 /* "Mat::operator(): index out of bounds" */
#define EUT_ORACLE_ERR_ALLOC 2

/* One Jacobi substep of one column, exactly in the reference's operation order
 * (src/diffChange.cpp:16-29 and :50-63):
 *   ychg = ychg * 0                        (NOT a fresh zero fill: keeps sign / NaN of old ychg)
 *   for every edge row j:   ychg[to_j] += col[from_j]
 *   ychg = ychg * (1/13)
 *   for every nneighbors row j: ychg[id_j] -= col[id_j] * ((nn_j - 1) * (1/13))
 *   col += ychg
 * from/to/id are already converted to 0-based int by the callers below.
 */
static int substep_column(double *col, double *ychg, size_t n_rows,
                          const int *from0, const int *to0, size_t n_edges,
                          const int *id0, const double *nn_as_double, size_t n_nn,
                          double b_div)
{
    for (size_t r = 0; r < n_rows; r++)
        ychg[r] = ychg[r] * 0;

    for (size_t j = 0; j < n_edges; j++) {
        int t = to0[j], f = from0[j];
        if (t < 0 || (size_t)t >= n_rows || f < 0 || (size_t)f >= n_rows)
            return EUT_ORACLE_ERR_INDEX;
        ychg[t] += col[f];
    }
    for (size_t r = 0; r < n_rows; r++)
        ychg[r] = ychg[r] * b_div;

    for (size_t j = 0; j < n_nn; j++) {
        int p = id0[j];
        if (p < 0 || (size_t)p >= n_rows)
            return EUT_ORACLE_ERR_INDEX;
        ychg[p] -= col[p] * ((nn_as_double[j] - 1) * b_div);
    }
    for (size_t r = 0; r < n_rows; r++)
        col[r] += ychg[r];
    return EUT_ORACLE_OK;
}

/* diffChange(adjmat, nneighbors, conc, niter)  -- src/diffChange.cpp:6-35.
 *   adjmat      n_edges x 2 doubles, column-major: [from | to]   (cast with (int), :20)
 *   nneighbors  n_nn x 2 doubles, column-major:    [id | n_neighbors(+1 for the diagonal)] (:26)
 *   conc        n_rows x n_cols doubles, column-major; updated IN PLACE here (the reference works
 *               on its by-value copy and returns it, :34 -- callers pass a copy)
 *   niter       n_niter doubles; the column loop runs over niter.size() (:12) and the substep
 *               loop is `for (int k = 0; k < niter(i); k++)` with niter(i) a double (:14)
 * ychg is allocated once, zero-filled, and shared by all columns (:10) -- kept that way.
 */
int eut_oracle_diffchange(const double *adjmat, size_t n_edges,
                          const double *nneighbors, size_t n_nn,
                          double *conc, size_t n_rows, size_t n_cols,
                          const double *niter, size_t n_niter)
{
    const double b_div = 1.0 / 13;
    int rc = EUT_ORACLE_OK;
    double *ychg = (double *)calloc(n_rows ? n_rows : 1, sizeof(double));
    int *from0 = (int *)malloc((n_edges ? n_edges : 1) * sizeof(int));
    int *to0 = (int *)malloc((n_edges ? n_edges : 1) * sizeof(int));
    int *id0 = (int *)malloc((n_nn ? n_nn : 1) * sizeof(int));
    if (!ychg || !from0 || !to0 || !id0) { rc = EUT_ORACLE_ERR_ALLOC; goto done; }

    for (size_t j = 0; j < n_edges; j++) {
        from0[j] = ((int)adjmat[j]) - 1;
        to0[j] = ((int)adjmat[n_edges + j]) - 1;
    }
    for (size_t j = 0; j < n_nn; j++)
        id0[j] = ((int)nneighbors[j]) - 1;

    for (size_t i = 0; i < n_niter; i++) {
        for (int k = 0; k < niter[i]; k++) {
            if (i >= n_cols) { rc = EUT_ORACLE_ERR_INDEX; goto done; } /* conc(.., i) out of range */
            rc = substep_column(conc + i * n_rows, ychg, n_rows, from0, to0, n_edges,
                                id0, nneighbors + n_nn, n_nn, b_div);
            if (rc) goto done;
        }
    }
done:
    free(ychg); free(from0); free(to0); free(id0);
    return rc;
}

/* diffChangePar(adjmat, nneighbors, conc, niter, indvar, ncores) -- src/diffChange.cpp:38-69.
 *   adjmat / nneighbors  int32, column-major, 1-based (:39)
 *   niter[i]             substeps of column indvar[i] (indexed by POSITION in indvar, :48)
 *   indvar               1-based column numbers (:54); must be unique for the OpenMP loop to be
 *                        race-free, exactly as in the reference (:44-45, :63)
 *   ychg                 thread-iteration private, fresh zeros per column (:47)
 */
int eut_oracle_diffchange_par(const int *adjmat, size_t n_edges,
                              const int *nneighbors, size_t n_nn,
                              double *conc, size_t n_rows, size_t n_cols,
                              const int *niter, const int *indvar, size_t n_indvar,
                              int ncores)
{
    const double b_div = 1.0 / 13;
    int rc_all = EUT_ORACLE_OK;
    if (ncores < 1) ncores = 1;

    int *from0 = (int *)malloc((n_edges ? n_edges : 1) * sizeof(int));
    int *to0 = (int *)malloc((n_edges ? n_edges : 1) * sizeof(int));
    int *id0 = (int *)malloc((n_nn ? n_nn : 1) * sizeof(int));
    double *nnd = (double *)malloc((n_nn ? n_nn : 1) * sizeof(double));
    if (!from0 || !to0 || !id0 || !nnd) {
        free(from0); free(to0); free(id0); free(nnd);
        return EUT_ORACLE_ERR_ALLOC;
    }
    for (size_t j = 0; j < n_edges; j++) {
        from0[j] = adjmat[j] - 1;
        to0[j] = adjmat[n_edges + j] - 1;
    }
    for (size_t j = 0; j < n_nn; j++) {
        id0[j] = nneighbors[j] - 1;
        nnd[j] = (double)nneighbors[n_nn + j];
    }

#pragma omp parallel for num_threads(ncores) schedule(static)
    for (long i = 0; i < (long)n_indvar; i++) {
        int rc = EUT_ORACLE_OK;
        double *ychg = (double *)calloc(n_rows ? n_rows : 1, sizeof(double));
        if (!ychg) rc = EUT_ORACLE_ERR_ALLOC;
        for (int k = 0; !rc && k < niter[i]; k++) {
            long c = (long)indvar[i] - 1;
            if (c < 0 || (size_t)c >= n_cols) { rc = EUT_ORACLE_ERR_INDEX; break; }
            rc = substep_column(conc + (size_t)c * n_rows, ychg, n_rows, from0, to0, n_edges,
                                id0, nnd, n_nn, b_div);
        }
        free(ychg);
        if (rc) {
#pragma omp critical
            rc_all = rc;
        }
    }
    free(from0); free(to0); free(id0); free(nnd);
    return rc_all;
}

int eut_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
