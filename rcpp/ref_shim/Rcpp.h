// TEST INFRASTRUCTURE ONLY -- a stand-in for <Rcpp.h>, written from Rcpp's documented behaviour (not
// copied), with just the facilities rcpp/diffChange.cpp and rcpp/session_glue.cpp use, so that the glue
// a maintainer drops into the R package meets a compiler and can be called in an image without R:
//   NumericMatrix / IntegerMatrix / NumericVector / IntegerVector   column-major, reference semantics
//                                                                    (copies share the data, like SEXPs)
//   stop(fmt, ...)              throws Rcpp::exception (what BEGIN_RCPP / END_RCPP turn into an R error)
//   XPtr<T, Storage, Finalizer> external pointer with tag / prot slots; the finalizer runs when the
//                               last handle goes away (R: when the GC collects the SEXP)
//   List::create(Named(..) = ..), SEXP, R_NilValue
// rcpp/shim_exports.cpp plays the part of the generated src/RcppExports.cpp.
#pragma once
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

struct SexpRec {                       // what a SEXP points at in this stand-in
    enum Kind { NIL, REAL, INT, EXTPTR, LIST } kind = NIL;
    std::vector<double> real;
    std::vector<int> integer;
    int nrow = 0, ncol = 0;            // ncol < 0: a plain vector
    std::shared_ptr<void> ext;         // external pointer payload (deleter = finalizer)
    std::shared_ptr<SexpRec> tag, prot;
    std::vector<std::pair<std::string, std::shared_ptr<SexpRec>>> items;
};
typedef std::shared_ptr<SexpRec> SEXP;
static const SEXP R_NilValue = SEXP();

namespace Rcpp {

class exception : public std::runtime_error {
public:
    explicit exception(const std::string &m) : std::runtime_error(m) {}
};

inline void stop(const char *fmt, ...)
{
    char buf[2048];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    throw exception(buf);
}

struct PreserveStorage {};

namespace detail {
template <typename T> struct store;
template <> struct store<double> {
    static std::vector<double> &of(SexpRec &r) { return r.real; }
    static constexpr SexpRec::Kind kind = SexpRec::REAL;
};
template <> struct store<int> {
    static std::vector<int> &of(SexpRec &r) { return r.integer; }
    static constexpr SexpRec::Kind kind = SexpRec::INT;
};
}  // namespace detail

template <typename T>
class Vector {
public:
    Vector() : s_(std::make_shared<SexpRec>()) { s_->kind = detail::store<T>::kind; s_->ncol = -1; }
    explicit Vector(long n) : Vector() { detail::store<T>::of(*s_).assign((size_t)n, T()); s_->nrow = (int)n; }
    Vector(const T *p, long n) : Vector() { detail::store<T>::of(*s_).assign(p, p + n); s_->nrow = (int)n; }
    Vector(SEXP s) : s_(std::move(s)) { if (!s_ || s_->kind != detail::store<T>::kind) stop("not compatible with requested type"); }
    T *begin() { return detail::store<T>::of(*s_).data(); }
    const T *begin() const { return detail::store<T>::of(*s_).data(); }
    long size() const { return (long)detail::store<T>::of(*s_).size(); }
    T &operator[](long i) { return detail::store<T>::of(*s_)[(size_t)i]; }
    operator SEXP() const { return s_; }
protected:
    SEXP s_;
};

template <typename T>
class Matrix : public Vector<T> {
public:
    Matrix(int nrow, int ncol) : Vector<T>((long)nrow * ncol) { this->s_->nrow = nrow; this->s_->ncol = ncol; }
    Matrix(const T *p, int nrow, int ncol) : Vector<T>(p, (long)nrow * ncol) { this->s_->nrow = nrow; this->s_->ncol = ncol; }
    Matrix(SEXP s) : Vector<T>(std::move(s)) { if (this->s_->ncol < 0) stop("not a matrix"); }
    int nrow() const { return this->s_->nrow; }
    int ncol() const { return this->s_->ncol; }
};
typedef Vector<double> NumericVector;
typedef Vector<int> IntegerVector;
typedef Matrix<double> NumericMatrix;
typedef Matrix<int> IntegerMatrix;

template <typename T, typename Storage, void (*Finalizer)(T *)>
class XPtr {
public:
    XPtr(T *p, bool set_delete_finalizer = true, SEXP tag = R_NilValue, SEXP prot = R_NilValue) : s_(std::make_shared<SexpRec>())
    {
        s_->kind = SexpRec::EXTPTR;
        if (set_delete_finalizer) s_->ext = std::shared_ptr<void>(p, [](void *q) { if (q) Finalizer(static_cast<T *>(q)); });
        else s_->ext = std::shared_ptr<void>(p, [](void *) {});
        s_->tag = std::move(tag);
        s_->prot = std::move(prot);          // kept alive as long as this pointer is (R_MakeExternalPtr's prot)
    }
    XPtr(SEXP s) : s_(std::move(s)) { if (!s_ || s_->kind != SexpRec::EXTPTR) stop("expecting an external pointer"); }
    T *get() const { return static_cast<T *>(s_->ext.get()); }
    operator SEXP() const { return s_; }
private:
    SEXP s_;
};

struct NamedValue { std::string name; SEXP value; };
struct Named {
    std::string name;
    explicit Named(const char *n) : name(n) {}
    template <typename V> NamedValue operator=(const V &v) const { return NamedValue{name, (SEXP)v}; }
};
class List {
public:
    template <typename... A> static List create(const A &...a) { List l; (void)std::initializer_list<int>{(l.s_->items.emplace_back(a.name, a.value), 0)...}; return l; }
    List() : s_(std::make_shared<SexpRec>()) { s_->kind = SexpRec::LIST; }
    List(SEXP s) : s_(std::move(s)) {}
    SEXP operator[](const std::string &n) const { for (auto &it : s_->items) if (it.first == n) return it.second; return R_NilValue; }
    operator SEXP() const { return s_; }
private:
    SEXP s_;
};

}  // namespace Rcpp
