"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import this package.  The product package `eutropia_b200` never does (tests/test_boundary.py
greps for it).

Parity pin: the reference ships no tests or golden vectors, but its one compiled source file,
src/diffChange.cpp, builds unmodified against a stand-in for the Armadillo types it uses
(`ref_shim/RcppArmadillo.h`) into `oracle/_ref/libeutropia_ref.so`; `ref_diffChange*` call it and
tests/test_oracle.py holds the restatement to it bit for bit.  The R-language parts of the path
(cell lists, gather, scatter, exoenzymes) have no such pin: R is absent -- "parity unpinned" there.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_eutropia.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (see oracle/Makefile)."""
    srcs = [os.path.join(_HERE, f) for f in ("diffchange_oracle.c", "cellfield_oracle.c", "exoenzyme_oracle.c", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        sz, dp, ip, lp = C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p
        L.eut_oracle_diffchange.argtypes = [dp, sz, dp, sz, dp, sz, sz, dp, sz]
        L.eut_oracle_diffchange.restype = C.c_int
        L.eut_oracle_diffchange_par.argtypes = [ip, sz, ip, sz, dp, sz, sz, ip, ip, sz, C.c_int]
        L.eut_oracle_diffchange_par.restype = C.c_int
        L.eut_oracle_max_threads.restype = C.c_int
        L.eut_oracle_cellmap.argtypes = [dp, sz, dp, dp, dp, dp, sz, C.c_int, lp, ip, dp, dp, C.c_long]
        L.eut_oracle_cellmap.restype = C.c_long
        L.eut_oracle_claim.argtypes = [ip, dp, C.c_long, sz]
        L.eut_oracle_claim.restype = C.c_int
        L.eut_oracle_gather.argtypes = [dp, sz, sz, lp, ip, dp, sz, C.c_double, dp, dp]
        L.eut_oracle_gather.restype = C.c_int
        L.eut_oracle_scatter.argtypes = [dp, sz, sz, C.c_void_p, lp, ip, dp, dp, sz, C.c_double,
                                         lp, ip, dp]
        L.eut_oracle_scatter.restype = C.c_int
        L.oracle_lambert_w0.argtypes = [C.c_double]
        L.oracle_lambert_w0.restype = C.c_double
        L.oracle_exoenzyme_catalysis.argtypes = [dp, C.c_int64, C.c_int64, dp, C.c_int64, C.c_int32, C.c_double,
                                                 C.c_double, ip, dp, C.c_int32, C.c_void_p, C.c_double]
        L.oracle_exoenzyme_catalysis.restype = C.c_int
        L.oracle_exoenzyme_decay.argtypes = [dp, C.c_int64, C.c_int32, C.c_double, C.c_double]
        L.oracle_exoenzyme_decay.restype = None
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, order="F"):
    return np.require(a, dtype=np.float64, requirements=[order, "A"])


class OracleIndexError(IndexError):
    """Mirrors the R error Armadillo's bounds check raises (`Mat::operator(): index out of bounds`)."""


def _check(rc):
    if rc == 1:
        raise OracleIndexError("Mat::operator(): index out of bounds")
    if rc:
        raise MemoryError("oracle allocation failed")


def max_threads() -> int:
    return int(lib().eut_oracle_max_threads())


def diffChange(adjmat, nneighbors, conc, niter):
    """`diffChange(adjmat, nneighbors, conc, niter)` (src/diffChange.cpp:6-35): every argument is
    converted to double like `arma::mat` / `arma::vec` do; returns a new matrix."""
    adj = _f64(np.asarray(adjmat).reshape(-1, 2) if np.size(adjmat) else np.zeros((0, 2)))
    nn = _f64(np.asarray(nneighbors).reshape(-1, 2) if np.size(nneighbors) else np.zeros((0, 2)))
    out = np.array(conc, dtype=np.float64, order="F", copy=True)
    if out.ndim == 1:
        out = out.reshape(-1, 1, order="F")
    nit = np.ascontiguousarray(np.asarray(niter, dtype=np.float64).ravel())
    rc = lib().eut_oracle_diffchange(_p(adj), adj.shape[0], _p(nn), nn.shape[0], _p(out),
                                     out.shape[0], out.shape[1], _p(nit), nit.shape[0])
    _check(rc)
    return out


def diffChangePar(adjmat, nneighbors, conc, niter, indvar, ncores=1):
    """`diffChangePar(...)` (src/diffChange.cpp:38-69): int32 indices (doubles are truncated like
    Rcpp's conversion to `arma::Col<int>`), OpenMP over `indvar`; returns a new matrix."""
    adj = np.require(np.asarray(adjmat).reshape(-1, 2), dtype=np.int32, requirements=["F", "A"])
    nn = np.require(np.asarray(nneighbors).reshape(-1, 2), dtype=np.int32, requirements=["F", "A"])
    out = np.array(conc, dtype=np.float64, order="F", copy=True)
    if out.ndim == 1:
        out = out.reshape(-1, 1, order="F")
    nit = np.ascontiguousarray(np.trunc(np.asarray(niter, dtype=np.float64)).astype(np.int32).ravel())
    iv = np.ascontiguousarray(np.trunc(np.asarray(indvar, dtype=np.float64)).astype(np.int32).ravel())
    if nit.shape[0] < iv.shape[0]:
        raise OracleIndexError("Mat::operator(): index out of bounds")  # niter(i), :48
    rc = lib().eut_oracle_diffchange_par(_p(adj), adj.shape[0], _p(nn), nn.shape[0], _p(out),
                                         out.shape[0], out.shape[1], _p(nit), _p(iv), iv.shape[0],
                                         int(ncores))
    _check(rc)
    return out


# ---- oracle/_ref: the reference's own src/diffChange.cpp, compiled unmodified (oracle/Makefile `ref`) ----
_REF_PATH = os.path.join(_HERE, "_ref", "libeutropia_ref.so")
_REF_SRC = "/root/reference/src/diffChange.cpp"
_ref = None


def ref_build(force: bool = False) -> bool:
    """Compile oracle/_ref when the reference tree is present (this container); on the GPU box the
    prebuilt library travels with the snapshot.  Returns whether the library exists afterwards."""
    deps = [os.path.join(_HERE, "ref_glue.cpp"), os.path.join(_HERE, "ref_shim", "RcppArmadillo.h")]
    if os.path.exists(_REF_SRC):
        stale = (not os.path.exists(_REF_PATH)) or any(
            os.path.getmtime(s) > os.path.getmtime(_REF_PATH) for s in deps + [_REF_SRC])
        if force or stale:
            subprocess.run(["make", "-C", _HERE, "-s", "ref"], check=True, stdout=subprocess.DEVNULL)
    return os.path.exists(_REF_PATH)


def ref_available() -> bool:
    """Whether oracle/_ref exists; compiles it only when it is missing and the reference tree is
    there (tests in the build container).  On the GPU box the tree is absent: prebuilt file or nothing."""
    return os.path.exists(_REF_PATH) or ref_build()


def ref_lib() -> C.CDLL:
    global _ref
    if _ref is None:
        if not ref_available():
            raise RuntimeError("oracle/_ref/libeutropia_ref.so is absent (built only where /root/reference exists)")
        L = C.CDLL(_REF_PATH)
        sz, vp = C.c_size_t, C.c_void_p
        L.eut_ref_diffchange.argtypes = [vp, sz, vp, sz, vp, sz, sz, vp, sz]
        L.eut_ref_diffchange.restype = C.c_int
        L.eut_ref_diffchange_par.argtypes = [vp, sz, vp, sz, vp, sz, sz, vp, sz, vp, sz, C.c_int]
        L.eut_ref_diffchange_par.restype = C.c_int
        L.eut_ref_max_threads.restype = C.c_int
        _ref = L
    return _ref


def _ref_check(rc):
    if rc == 1:
        raise OracleIndexError("Mat::operator(): index out of bounds")
    if rc:
        raise RuntimeError("reference library raised")


def ref_diffChange(adjmat, nneighbors, conc, niter):
    """The reference's own `diffChange` (src/diffChange.cpp:6-35) through oracle/_ref."""
    adj = _f64(np.asarray(adjmat).reshape(-1, 2) if np.size(adjmat) else np.zeros((0, 2)))
    nn = _f64(np.asarray(nneighbors).reshape(-1, 2) if np.size(nneighbors) else np.zeros((0, 2)))
    out = np.array(conc, dtype=np.float64, order="F", copy=True)
    if out.ndim == 1:
        out = out.reshape(-1, 1, order="F")
    nit = np.ascontiguousarray(np.asarray(niter, dtype=np.float64).ravel())
    _ref_check(ref_lib().eut_ref_diffchange(_p(adj), adj.shape[0], _p(nn), nn.shape[0], _p(out),
                                            out.shape[0], out.shape[1], _p(nit), nit.shape[0]))
    return out


def ref_diffChangePar(adjmat, nneighbors, conc, niter, indvar, ncores=1):
    """The reference's own `diffChangePar` (src/diffChange.cpp:38-69) through oracle/_ref.  An index
    error inside its OpenMP region would terminate the process (as it does under R), so callers
    pass valid indices only."""
    adj = np.require(np.asarray(adjmat).reshape(-1, 2), dtype=np.int32, requirements=["F", "A"])
    nn = np.require(np.asarray(nneighbors).reshape(-1, 2), dtype=np.int32, requirements=["F", "A"])
    out = np.array(conc, dtype=np.float64, order="F", copy=True)
    if out.ndim == 1:
        out = out.reshape(-1, 1, order="F")
    nit = np.ascontiguousarray(np.trunc(np.asarray(niter, dtype=np.float64)).astype(np.int32).ravel())
    iv = np.ascontiguousarray(np.trunc(np.asarray(indvar, dtype=np.float64)).astype(np.int32).ravel())
    n, c = out.shape
    if nit.shape[0] < iv.shape[0] or (iv.size and (iv.min() < 1 or iv.max() > c)) or \
            (adj.size and (adj.min() < 1 or adj.max() > n)) or (nn.size and (nn[:, 0].min() < 1 or nn[:, 0].max() > n)):
        raise OracleIndexError("Mat::operator(): index out of bounds")
    _ref_check(ref_lib().eut_ref_diffchange_par(_p(adj), adj.shape[0], _p(nn), nn.shape[0], _p(out), n, c,
                                                _p(nit), nit.shape[0], _p(iv), iv.shape[0], int(ncores)))
    return out


def ref_max_threads() -> int:
    return int(ref_lib().eut_ref_max_threads())


def cellmap(pts, cx, cy, csize, cscv, dist_mode=0):
    """Cell -> field lists; returns (cell_ptr int64[M+1], field_id int32, field_dist, acc_prop)."""
    pts = _f64(pts)
    cx, cy, csize, cscv = (np.ascontiguousarray(a, dtype=np.float64) for a in (cx, cy, csize, cscv))
    m = cx.shape[0]
    ptr = np.zeros(m + 1, dtype=np.int64)
    cap = max(1, 256 * m)
    while True:
        fid = np.zeros(cap, dtype=np.int32)
        fd = np.zeros(cap, dtype=np.float64)
        acc = np.zeros(cap, dtype=np.float64)
        tot = lib().eut_oracle_cellmap(_p(pts), pts.shape[0], _p(cx), _p(cy), _p(csize), _p(cscv), m,
                                       int(dist_mode), _p(ptr), _p(fid), _p(fd), _p(acc), cap)
        if tot < 0:
            raise MemoryError
        if tot <= cap:
            return ptr, fid[:tot].copy(), fd[:tot].copy(), acc[:tot].copy()
        cap = int(tot)


def claim(field_id, acc_prop, n_pts):
    """Oversubscription rescale; returns the new acc_prop."""
    fid = np.ascontiguousarray(field_id, dtype=np.int32)
    acc = np.array(acc_prop, dtype=np.float64, copy=True)
    _check(lib().eut_oracle_claim(_p(fid), _p(acc), fid.shape[0], int(n_pts)))
    return acc


def gather(conc, cell_ptr, field_id, acc_prop, field_vol, want_fpf=False):
    conc = _f64(conc)
    ptr = np.ascontiguousarray(cell_ptr, dtype=np.int64)
    fid = np.ascontiguousarray(field_id, dtype=np.int32)
    acc = np.ascontiguousarray(acc_prop, dtype=np.float64)
    m = ptr.shape[0] - 1
    fmol = np.zeros((m, conc.shape[1]), dtype=np.float64, order="F")
    fpf = np.zeros((fid.shape[0], conc.shape[1]), dtype=np.float64, order="F") if want_fpf else None
    _check(lib().eut_oracle_gather(_p(conc), conc.shape[0], conc.shape[1], _p(ptr), _p(fid), _p(acc),
                                   m, float(field_vol), _p(fmol), _p(fpf)))
    return (fmol, fpf) if want_fpf else fmol


def scatter(conc, is_constant, cell_ptr, field_id, field_dist, acc_prop, field_vol,
            ex_ptr, ex_cpd, ex_flx):
    out = np.array(conc, dtype=np.float64, order="F", copy=True)
    isc = None if is_constant is None else np.ascontiguousarray(is_constant, dtype=np.uint8)
    ptr = np.ascontiguousarray(cell_ptr, dtype=np.int64)
    fid = np.ascontiguousarray(field_id, dtype=np.int32)
    fd = np.ascontiguousarray(field_dist, dtype=np.float64)
    acc = np.ascontiguousarray(acc_prop, dtype=np.float64)
    exp_ = np.ascontiguousarray(ex_ptr, dtype=np.int64)
    exc = np.ascontiguousarray(ex_cpd, dtype=np.int32)
    exf = np.ascontiguousarray(ex_flx, dtype=np.float64)
    _check(lib().eut_oracle_scatter(_p(out), out.shape[0], out.shape[1], _p(isc), _p(ptr), _p(fid),
                                    _p(fd), _p(acc), ptr.shape[0] - 1, float(field_vol),
                                    _p(exp_), _p(exc), _p(exf)))
    return out


def lambertW0(x):
    """Principal branch of the Lambert W function (lamW::lambertW0), long double Halley iteration."""
    x = np.asarray(x, dtype=np.float64)
    f = lib().oracle_lambert_w0
    return np.vectorize(lambda v: f(float(v)), otypes=[np.float64])(x)


def exoenzyme_catalysis(conc, exo_conc, k, kcat, km, mets, stoich, is_constant, delta_time):
    """R/run_simulation.R:343-364 for exoenzyme `k` (1-based column of exo_conc); returns the new
    concentration matrix."""
    out = np.array(conc, dtype=np.float64, order="F", copy=True)
    exo = _f64(exo_conc)
    mets = np.ascontiguousarray(mets, dtype=np.int32)
    st = np.ascontiguousarray(stoich, dtype=np.float64)
    isc = None if is_constant is None else np.ascontiguousarray(is_constant, dtype=np.uint8)
    _check(lib().oracle_exoenzyme_catalysis(_p(out), out.shape[0], out.shape[1], _p(exo), exo.shape[1],
                                            int(k), float(kcat), float(km), _p(mets), _p(st),
                                            mets.shape[0], _p(isc), float(delta_time)))
    return out


def exoenzyme_decay(exo_conc, k, lam, delta_time):
    """R/run_simulation.R:367-371 for column `k` (1-based); returns the new exoenzyme matrix."""
    out = np.array(exo_conc, dtype=np.float64, order="F", copy=True)
    lib().oracle_exoenzyme_decay(_p(out), out.shape[0], int(k), float(lam), float(delta_time))
    return out


def chemotaxis_shift(conc_col, pts, cell_ptr, field_id, cell_x, cell_y, cell_size, scavenge_dist, hill_ka, hill_coef,
                     strength, vmax, delta_time):
    """R/run_simulation.R:260-291, cell by cell and statement by statement (numpy; small cases)."""
    m = len(cell_x)
    ctx, cty = np.zeros(m), np.zeros(m)
    for i in range(m):
        fid = np.asarray(field_id[cell_ptr[i]:cell_ptr[i + 1]], dtype=np.int64) - 1
        r = cell_size[i] / 2 + scavenge_dist
        dx = (pts[fid, 0] - cell_x[i]) / r
        dy = (pts[fid, 1] - cell_y[i]) / r
        f_orig = conc_col[fid]
        f = f_orig.copy()
        if f.sum() == 0:
            f = np.ones_like(f)
        f = f / f.sum()
        cx, cy = (dx * f).sum() / f.sum(), (dy * f).sum() / f.sum()
        with np.errstate(divide="ignore"):
            sens = 1 / (1 + (hill_ka / ((f_orig * f).sum() / f.sum())) ** hill_coef)
        ctx[i] = cx * sens * strength * vmax * 60 * delta_time * 60
        cty[i] = cy * sens * strength * vmax * 60 * delta_time * 60
    return ctx, cty
