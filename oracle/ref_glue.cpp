// TEST INFRASTRUCTURE ONLY.  C entry points around the reference's own diffChange / diffChangePar
// (src/diffChange.cpp:6 and :38, compiled unmodified into the same library, see oracle/Makefile
// target `ref`).  They do what src/RcppExports.cpp:16-54 does for R: wrap the caller's column-major
// buffers into arma objects, call, copy the returned matrix out.
#include <RcppArmadillo.h>

#include <cstdint>
#include <exception>
#ifdef _OPENMP
#include <omp.h>
#endif

arma::mat diffChange(arma::mat adjmat, arma::mat nneighbors, arma::mat conc, arma::vec niter);
arma::mat diffChangePar(arma::Mat<int> adjmat, arma::Mat<int> nneighbors, arma::Mat<double> conc,
                        arma::Col<int> niter, arma::Col<int> indvar, int ncores);

extern "C" {

// 0 ok, 1 Armadillo bounds error, 2 anything else
int eut_ref_diffchange(const double *adj, size_t n_edges, const double *nn, size_t n_nn, double *conc,
                       size_t n, size_t n_cols, const double *niter, size_t n_iter)
{
    try {
        arma::mat out = diffChange(arma::mat(adj, n_edges, 2), arma::mat(nn, n_nn, 2), arma::mat(conc, n, n_cols),
                                   arma::vec(niter, n_iter));
        std::memcpy(conc, out.memptr(), sizeof(double) * n * n_cols);
        return 0;
    } catch (const std::out_of_range &) {
        return 1;
    } catch (const std::exception &) {
        return 2;
    }
}

int eut_ref_diffchange_par(const int *adj, size_t n_edges, const int *nn, size_t n_nn, double *conc, size_t n,
                           size_t n_cols, const int *niter, size_t n_iter, const int *indvar, size_t n_ind,
                           int ncores)
{
    try {
        arma::mat out = diffChangePar(arma::Mat<int>(adj, n_edges, 2), arma::Mat<int>(nn, n_nn, 2),
                                      arma::mat(conc, n, n_cols), arma::Col<int>(niter, n_iter),
                                      arma::Col<int>(indvar, n_ind), ncores);
        std::memcpy(conc, out.memptr(), sizeof(double) * n * n_cols);
        return 0;
    } catch (const std::out_of_range &) {
        return 1;
    } catch (const std::exception &) {
        return 2;
    }
}

int eut_ref_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
