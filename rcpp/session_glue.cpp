// Optional second half of the R integration: the resident session.  With only rcpp/diffChange.cpp
// every round still moves the whole N x C matrix over PCIe twice (the price of R's value semantics
// at the diffChangePar call, SURVEY 8b "Copy semantics").  These exports let a patched
// diffuse_compoundsPar / run_simulation keep `concentrations` on the GPU between rounds and pull
// only what R reads.  External pointers own the C handles; R's GC releases them.
//
// Compiled here against the stand-in rcpp/ref_shim/Rcpp.h (rcpp/Makefile); with R it builds against the real Rcpp.
#include <Rcpp.h>

#include <eutropia_b200.h>

namespace {
void check(int rc) { if (rc != EUT_OK) Rcpp::stop("eutropia_b200: %s", eut_last_error()); }
void free_plan(eut_plan *p) { eut_plan_destroy(p); }
void free_field(eut_field *f) { eut_field_destroy(f); }
void free_cells(eut_cells *c) { eut_cells_destroy(c); }
typedef Rcpp::XPtr<eut_plan, Rcpp::PreserveStorage, free_plan> PlanPtr;
typedef Rcpp::XPtr<eut_field, Rcpp::PreserveStorage, free_field> FieldPtr;
typedef Rcpp::XPtr<eut_cells, Rcpp::PreserveStorage, free_cells> CellsPtr;
}  // namespace

// [[Rcpp::export]]
SEXP eutb_plan(Rcpp::IntegerMatrix mat_in, Rcpp::IntegerMatrix mat_out, int device = 0)
{
    eut_plan *p = nullptr;   // mat.in = (from | to), mat.out = (id | n_neighbors): R/growthEnvironment.R:145-146
    check(eut_plan_create(&p, mat_in.begin(), mat_in.begin() + mat_in.nrow(), mat_in.nrow(), mat_out.begin(),
                          mat_out.begin() + mat_out.nrow(), mat_out.nrow(), mat_out.nrow(), device, EUT_ORDER_AUTO));
    return PlanPtr(p, true);
}

// [[Rcpp::export]]
SEXP eutb_field(SEXP plan, Rcpp::NumericMatrix conc)
{
    eut_field *f = nullptr;
    check(eut_field_create(&f, PlanPtr(plan).get(), conc.ncol()));
    const int rc = eut_field_upload(f, conc.begin(), conc.nrow(), nullptr, conc.ncol());
    if (rc != EUT_OK) { eut_field_destroy(f); check(rc); }
    // the field refers to the plan: the plan's external pointer sits in the prot slot, so R cannot
    // finalise it before the field (eut_field_destroy dereferences the plan)
    return FieldPtr(f, true, R_NilValue, plan);
}

// [[Rcpp::export]]
void eutb_diffuse(SEXP field, Rcpp::IntegerVector niter, Rcpp::IntegerVector indvar)
{
    if (niter.size() < indvar.size()) Rcpp::stop("Mat::operator(): index out of bounds");   // niter(i), src/diffChange.cpp:48
    check(eut_field_diffuse(FieldPtr(field).get(), niter.begin(), indvar.begin(), indvar.size()));
}

// [[Rcpp::export]]
Rcpp::NumericMatrix eutb_download(SEXP field, int nrow, Rcpp::IntegerVector cols)
{
    Rcpp::NumericMatrix out(nrow, cols.size());
    check(eut_field_download(FieldPtr(field).get(), out.begin(), nrow, cols.begin(), cols.size()));
    return out;
}

// [[Rcpp::export]]
Rcpp::List eutb_column_stats(SEXP field, int ncol)
{
    Rcpp::NumericVector mn(ncol), mx(ncol), mean(ncol);   // sd > 0  <=>  max > min (R/growthEnvironment.R:372)
    check(eut_field_column_stats(FieldPtr(field).get(), mn.begin(), mx.begin(), mean.begin()));
    return Rcpp::List::create(Rcpp::Named("min") = mn, Rcpp::Named("max") = mx, Rcpp::Named("mean") = mean);
}
