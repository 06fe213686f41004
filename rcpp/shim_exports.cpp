// TEST INFRASTRUCTURE ONLY.  Plays the part of the generated src/RcppExports.cpp (reference lines 16-54)
// for an image without R: C entry points that wrap the caller's column-major buffers into (stand-in)
// Rcpp objects, call the exported glue functions of rcpp/diffChange.cpp and rcpp/session_glue.cpp
// exactly as `.Call` would, copy the result out and turn a C++ exception into a status + message --
// what BEGIN_RCPP / END_RCPP do with an R error.  Bound by tests/test_rcpp_glue.py through ctypes.
#include <Rcpp.h>

#include <cstdint>
#include <cstring>
#include <string>

Rcpp::NumericMatrix diffChange(Rcpp::NumericMatrix adjmat, Rcpp::NumericMatrix nneighbors, Rcpp::NumericMatrix conc,
                               Rcpp::NumericVector niter);
Rcpp::NumericMatrix diffChangePar(Rcpp::IntegerMatrix adjmat, Rcpp::IntegerMatrix nneighbors, Rcpp::NumericMatrix conc,
                                  Rcpp::IntegerVector niter, Rcpp::IntegerVector indvar, int ncores);
SEXP eutb_plan(Rcpp::IntegerMatrix mat_in, Rcpp::IntegerMatrix mat_out, int device);
SEXP eutb_field(SEXP plan, Rcpp::NumericMatrix conc);
void eutb_diffuse(SEXP field, Rcpp::IntegerVector niter, Rcpp::IntegerVector indvar);
Rcpp::NumericMatrix eutb_download(SEXP field, int nrow, Rcpp::IntegerVector cols);
Rcpp::List eutb_column_stats(SEXP field, int ncol);

namespace {
thread_local std::string g_msg;
template <typename F>
int guarded(F &&f)
{
    try { f(); g_msg.clear(); return 0; }
    catch (const Rcpp::exception &e) { g_msg = e.what(); return 1; }       // -> R error
    catch (const std::exception &e) { g_msg = e.what(); return 2; }
}
}  // namespace

#define GLUE_API extern "C" __attribute__((visibility("default")))

GLUE_API const char *glue_last_error(void) { return g_msg.c_str(); }

// .Call(`_Eutropia_diffChange`, adjmat, nneighbors, conc, niter)   src/RcppExports.cpp:16-27
GLUE_API int glue_Eutropia_diffChange(const double *adj, int n_edges, int adj_cols, const double *nn, int n_nn, int nn_cols,
                                      const double *conc, int n_rows, int n_cols, const double *niter, int n_niter, double *out)
{
    return guarded([&] {
        Rcpp::NumericMatrix r = diffChange(Rcpp::NumericMatrix(adj, n_edges, adj_cols), Rcpp::NumericMatrix(nn, n_nn, nn_cols),
                                           Rcpp::NumericMatrix(conc, n_rows, n_cols), Rcpp::NumericVector(niter, n_niter));
        std::memcpy(out, r.begin(), sizeof(double) * (size_t)n_rows * n_cols);
    });
}

// .Call(`_Eutropia_diffChangePar`, adjmat, nneighbors, conc, niter, indvar, ncores)   src/RcppExports.cpp:30-43
GLUE_API int glue_Eutropia_diffChangePar(const int *adj, int n_edges, int adj_cols, const int *nn, int n_nn, int nn_cols,
                                         const double *conc, int n_rows, int n_cols, const int *niter, int n_niter,
                                         const int *indvar, int n_indvar, int ncores, double *out)
{
    return guarded([&] {
        Rcpp::NumericMatrix r = diffChangePar(Rcpp::IntegerMatrix(adj, n_edges, adj_cols), Rcpp::IntegerMatrix(nn, n_nn, nn_cols),
                                              Rcpp::NumericMatrix(conc, n_rows, n_cols), Rcpp::IntegerVector(niter, n_niter),
                                              Rcpp::IntegerVector(indvar, n_indvar), ncores);
        std::memcpy(out, r.begin(), sizeof(double) * (size_t)n_rows * n_cols);
    });
}

// the resident session of rcpp/session_glue.cpp: plan -> field -> diffuse -> stats -> download, then both
// external pointers are dropped FIELD LAST's owner FIRST (the plan handle goes away before the field's:
// the prot slot must keep the plan alive until the field's finalizer has run)
GLUE_API int glue_session_round(const int *mat_in, int n_edges, const int *mat_out, int n, const double *conc, int n_cols,
                                const int *niter, const int *indvar, int n_indvar, double *out, double *col_max)
{
    return guarded([&] {
        SEXP field;
        {
            SEXP plan = eutb_plan(Rcpp::IntegerMatrix(mat_in, n_edges, 2), Rcpp::IntegerMatrix(mat_out, n, 2), 0);
            field = eutb_field(plan, Rcpp::NumericMatrix(conc, n, n_cols));
        }                                                    // R drops its reference to the plan here
        eutb_diffuse(field, Rcpp::IntegerVector(niter, n_indvar), Rcpp::IntegerVector(indvar, n_indvar));
        Rcpp::List st = eutb_column_stats(field, n_cols);
        Rcpp::NumericVector mx(st["max"]);
        std::memcpy(col_max, mx.begin(), sizeof(double) * (size_t)n_cols);
        Rcpp::IntegerVector cols(n_cols);
        for (int c = 0; c < n_cols; c++) cols[c] = c + 1;
        Rcpp::NumericMatrix r = eutb_download(field, n, cols);
        std::memcpy(out, r.begin(), sizeof(double) * (size_t)n * n_cols);
    });
}
