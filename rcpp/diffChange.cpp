// Drop-in replacement for the reference's src/diffChange.cpp: same two exported functions, same
// signatures, same R-level behaviour (value semantics, errors as R errors), but the work is done by
// libeutropia_b200.so on a B200 through its C ABI (include/eutropia_b200.h).
//
// This is the glue a maintainer drops into the R package next to the generated RcppExports.cpp, which
// needs NO change: the exported symbols _Eutropia_diffChange / _Eutropia_diffChangePar and
// R_init_Eutropia come out of Rcpp::compileAttributes() exactly as before (reference:
// src/RcppExports.cpp:16-54).  R and Rcpp are absent from this repository's image; here the file is
// compiled against the stand-in rcpp/ref_shim/Rcpp.h (rcpp/Makefile) and called by the tests through
// rcpp/shim_exports.cpp, which plays the part of RcppExports.cpp.
//
// Plain Rcpp matrices are used instead of Armadillo: nothing here does linear algebra, and this way
// the package no longer needs RcppArmadillo at all.
#include <Rcpp.h>

#include <cstdlib>
#include <vector>

#include <eutropia_b200.h>

namespace {
int eut_device()
{
    // one CUDA context per R process (PSOCK workers may call diffChange concurrently,
    // R/growthEnvironment.R:319); the ordinal can be pinned per worker
    const char *e = std::getenv("EUTROPIA_B200_DEVICE");
    return e ? std::atoi(e) : 0;
}
// diffChangePar is called once per round from the main R process and spreads its columns over OpenMP
// threads there (src/diffChange.cpp:44-45, R/growthEnvironment.R:385-390); here it spreads them over the
// GPUs of the box: EUTROPIA_B200_DEVICES = "0,2,3" picks them, unset = every visible device.
std::vector<int> eut_devices()
{
    std::vector<int> d;
    if (const char *e = std::getenv("EUTROPIA_B200_DEVICES"))
        for (const char *q = e; *q;) {
            char *end = nullptr;
            const long v = std::strtol(q, &end, 10);
            if (end == q) break;
            d.push_back((int)v);
            q = (*end == ',') ? end + 1 : end;
        }
    return d;
}
void check(int rc)
{
    if (rc != EUT_OK) Rcpp::stop("eutropia_b200: %s", eut_last_error());  // -> R error, like BEGIN_RCPP/END_RCPP
}
}  // namespace

// [[Rcpp::export]]
Rcpp::NumericMatrix diffChange(Rcpp::NumericMatrix adjmat, Rcpp::NumericMatrix nneighbors,
                               Rcpp::NumericMatrix conc, Rcpp::NumericVector niter)
{
    // NumericMatrix coerces integer R matrices to double, as arma::mat did (src/RcppExports.cpp:20-23)
    if (adjmat.ncol() != 2 || nneighbors.ncol() != 2) Rcpp::stop("adjmat / nneighbors must have two columns");
    Rcpp::NumericMatrix out(conc.nrow(), conc.ncol());   // fresh REALSXP: inputs are never written
    check(eut_diffchange(adjmat.begin(), adjmat.nrow(), nneighbors.begin(), nneighbors.nrow(),
                         conc.begin(), conc.nrow(), conc.ncol(), niter.begin(), niter.size(),
                         out.begin(), eut_device()));
    return out;
}

// [[Rcpp::export]]
Rcpp::NumericMatrix diffChangePar(Rcpp::IntegerMatrix adjmat, Rcpp::IntegerMatrix nneighbors,
                                  Rcpp::NumericMatrix conc, Rcpp::IntegerVector niter,
                                  Rcpp::IntegerVector indvar, int ncores)
{
    // IntegerVector truncates the doubles R passes for niter (ceiling(...), R/growthEnvironment.R:382)
    // exactly like arma::Col<int> did (src/RcppExports.cpp:37-38)
    if (adjmat.ncol() != 2 || nneighbors.ncol() != 2) Rcpp::stop("adjmat / nneighbors must have two columns");
    if (niter.size() < indvar.size()) Rcpp::stop("Mat::operator(): index out of bounds");   // niter(i), src/diffChange.cpp:48
    Rcpp::NumericMatrix out(conc.nrow(), conc.ncol());
    const std::vector<int> devs = eut_devices();
    check(eut_diffchange_par_multi(adjmat.begin(), adjmat.nrow(), nneighbors.begin(), nneighbors.nrow(),
                                   conc.begin(), conc.nrow(), conc.ncol(), niter.begin(), indvar.begin(),
                                   indvar.size(), ncores, out.begin(), devs.empty() ? nullptr : devs.data(),
                                   (int)devs.size()));
    return out;
}
