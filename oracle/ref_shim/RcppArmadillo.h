// TEST INFRASTRUCTURE ONLY.  Stand-in for <RcppArmadillo.h> so that the reference's own
// src/diffChange.cpp compiles UNMODIFIED, from where it lies, into oracle/_ref/ (see oracle/Makefile,
// target `ref`).  R, Rcpp and Armadillo are absent from this image; this header restates the small
// subset of the Armadillo 14 API that src/diffChange.cpp:1-69 touches, with Armadillo's documented
// semantics:
//   * Mat<T> / Col<T> are dense, column-major, value types (copy on pass-by-value);
//   * operator()(i) / operator()(i,j) are bounds checked (the reference does not define
//     ARMA_NO_DEBUG, src/Makevars) and throw std::out_of_range with Armadillo's message;
//   * `v * s` is the element-wise product with the scalar converted to the element type
//     (eop_scalar_times), evaluated element by element: `ychg * 0` keeps NaN and the sign of zero;
//   * `M.col(j) += v` adds element by element (subview_col::operator+=), sizes must agree.
// Nothing here is copied from Armadillo; only the names and meanings are the same.
#pragma once
#include <cstddef>
#include <cstring>
#include <stdexcept>
#include <vector>

namespace arma {

typedef unsigned long long uword;

namespace fill {
struct fill_zeros {};
static const fill_zeros zeros = fill_zeros();
}  // namespace fill

template <class T> class Col;

template <class T> class subview_col {
public:
    T *mem;
    uword n_rows;
    subview_col(T *m, uword n) : mem(m), n_rows(n) {}
    void operator+=(const Col<T> &v);
};

template <class T> class Mat {
public:
    uword n_rows, n_cols, n_elem;

    Mat() : n_rows(0), n_cols(0), n_elem(0) {}
    Mat(uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), store(r * c) {}
    Mat(const T *src, uword r, uword c) : n_rows(r), n_cols(c), n_elem(r * c), store(src, src + r * c) {}

    T *memptr() { return store.data(); }
    const T *memptr() const { return store.data(); }
    uword size() const { return n_elem; }

    T &operator()(uword i)
    {
        if (i >= n_elem) throw std::out_of_range("Mat::operator(): index out of bounds");
        return store[i];
    }
    const T &operator()(uword i) const
    {
        if (i >= n_elem) throw std::out_of_range("Mat::operator(): index out of bounds");
        return store[i];
    }
    T &operator()(uword i, uword j)
    {
        if (i >= n_rows || j >= n_cols) throw std::out_of_range("Mat::operator(): index out of bounds");
        return store[i + j * n_rows];
    }
    const T &operator()(uword i, uword j) const
    {
        if (i >= n_rows || j >= n_cols) throw std::out_of_range("Mat::operator(): index out of bounds");
        return store[i + j * n_rows];
    }
    subview_col<T> col(uword j)
    {
        if (j >= n_cols) throw std::out_of_range("Mat::col(): index out of bounds");
        return subview_col<T>(store.data() + j * n_rows, n_rows);
    }

protected:
    std::vector<T> store;
};

template <class T> class Col : public Mat<T> {
public:
    Col() : Mat<T>() {}
    explicit Col(uword n) : Mat<T>(n, 1) {}
    Col(uword n, const fill::fill_zeros &) : Mat<T>(n, 1) {}   // std::vector value-initialises: +0
    Col(const T *src, uword n) : Mat<T>(src, n, 1) {}
};

template <class T, class S> Col<T> operator*(const Col<T> &a, S s)
{
    Col<T> out(a.n_elem);
    const T k = (T)s;
    const T *x = a.memptr();
    T *y = out.memptr();
    for (uword i = 0; i < a.n_elem; i++) y[i] = x[i] * k;
    return out;
}

template <class T> void subview_col<T>::operator+=(const Col<T> &v)
{
    if (v.n_elem != n_rows) throw std::logic_error("addition: incompatible matrix dimensions");
    const T *x = v.memptr();
    for (uword i = 0; i < n_rows; i++) mem[i] += x[i];
}

typedef Mat<double> mat;
typedef Col<double> vec;

}  // namespace arma
