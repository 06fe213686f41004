"""The two Rcpp entry points of the hot path, same names / argument meaning / error behaviour
(R/RcppExports.R:4-10), bound to the one-shot tier of the C ABI.

The reference's toolchain (R, Rcpp, Armadillo) is absent from this image, so this module is the
host-side mirror that the parity tests call; the Rcpp glue that a maintainer drops into the R package
is under rcpp/ and described in INTEGRATION.md.  Both run exactly the same C entry points.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from ._lib import check, ptr

__all__ = ["diffChange", "diffChangePar"]


def _matrix(conc):
    a = np.asarray(conc)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.ndim != 2:
        raise ValueError("conc must be a matrix")
    return np.require(a, dtype=np.float64, requirements=["F", "A"])


def diffChange(adjmat, nneighbors, conc, niter, device: int = 0):
    """`diffChange(adjmat, nneighbors, conc, niter)` -- src/diffChange.cpp:6-35.

    Arguments are converted to double matrices / a double vector exactly like the generated glue
    does (`arma::mat`, `arma::vec`; src/RcppExports.cpp:20-23).  Returns a new N x C matrix; the
    inputs are never written (R value semantics)."""
    adj = np.require(np.asarray(adjmat, dtype=np.float64).reshape(-1, 2), requirements=["F", "A"])
    nn = np.require(np.asarray(nneighbors, dtype=np.float64).reshape(-1, 2), requirements=["F", "A"])
    c = _matrix(conc)
    nit = np.ascontiguousarray(np.asarray(niter, dtype=np.float64).ravel())
    out = np.empty(c.shape, dtype=np.float64, order="F")
    check(_lib.lib().eut_diffchange(ptr(adj), adj.shape[0], ptr(nn), nn.shape[0], ptr(c), c.shape[0],
                                    c.shape[1], ptr(nit), nit.shape[0], ptr(out), int(device)))
    return out


def diffChangePar(adjmat, nneighbors, conc, niter, indvar, ncores=1, device: int = 0, devices=None):
    """`diffChangePar(adjmat, nneighbors, conc, niter, indvar, ncores)` -- src/diffChange.cpp:38-69.

    `niter` / `indvar` may be doubles (R passes `ceiling(...)` results, R/growthEnvironment.R:382):
    they are truncated like Rcpp's conversion to `arma::Col<int>` (src/RcppExports.cpp:37-38).
    `ncores` is accepted and ignored.  `devices`: a list of CUDA ordinals (or "all") runs the call on
    several GPUs at once through `eut_diffchange_par_multi` -- contiguous column blocks per GPU, one
    result matrix -- the way the reference spreads the columns of one call over OpenMP threads
    (src/diffChange.cpp:44-45)."""
    adj = np.require(np.asarray(adjmat).reshape(-1, 2), dtype=np.int32, requirements=["F", "A"])
    nn = np.require(np.asarray(nneighbors).reshape(-1, 2), dtype=np.int32, requirements=["F", "A"])
    c = _matrix(conc)
    nit = np.ascontiguousarray(np.trunc(np.asarray(niter, dtype=np.float64)).astype(np.int32).ravel())
    iv = np.ascontiguousarray(np.trunc(np.asarray(indvar, dtype=np.float64)).astype(np.int32).ravel())
    if nit.shape[0] < iv.shape[0]:
        # niter(i) with i >= niter.size(): Armadillo's bounds check (src/diffChange.cpp:48)
        raise _lib.EutropiaIndexError(_lib.ERR_INDEX, "Mat::operator(): index out of bounds (niter)")
    out = np.empty(c.shape, dtype=np.float64, order="F")
    if devices is not None:
        dv = None if isinstance(devices, str) else np.ascontiguousarray(devices, dtype=np.int32)
        check(_lib.lib().eut_diffchange_par_multi(ptr(adj), adj.shape[0], ptr(nn), nn.shape[0], ptr(c), c.shape[0],
                                                  c.shape[1], ptr(nit), ptr(iv), iv.shape[0], int(ncores),
                                                  ptr(out), ptr(dv), 0 if dv is None else dv.shape[0]))
        return out
    check(_lib.lib().eut_diffchange_par(ptr(adj), adj.shape[0], ptr(nn), nn.shape[0], ptr(c), c.shape[0],
                                        c.shape[1], ptr(nit), ptr(iv), iv.shape[0], int(ncores),
                                        ptr(out), int(device)))
    return out
