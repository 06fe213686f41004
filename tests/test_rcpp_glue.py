"""The Rcpp glue of rcpp/ (what a maintainer drops into the R package) compiled against the stand-in
rcpp/ref_shim/Rcpp.h and called the way `.Call` would call the generated RcppExports.cpp
(src/RcppExports.cpp:16-54): rcpp/shim_exports.cpp wraps plain buffers into Rcpp objects, calls the two
exported functions, copies the result out and maps Rcpp::stop to a status + message.
CPU part: the library builds, loads, exports its symbols and raises the R-level errors that need no GPU.
GPU part: the reference's golden vectors through `diffChange` / `diffChangePar`, and the resident session
exports of rcpp/session_glue.cpp (with the plan's handle dropped before the field's, as R's GC may)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
LIB = os.path.join(ROOT, "rcpp", "_build", "libeutropia_rcpp_glue.so")


@pytest.fixture(scope="module")
def glue():
    if not os.path.exists(LIB):
        subprocess.run(["make", "-C", os.path.join(ROOT, "rcpp")], check=True)
    L = C.CDLL(LIB)
    vp, i = C.c_void_p, C.c_int
    L.glue_last_error.restype = C.c_char_p
    L.glue_Eutropia_diffChange.argtypes = [vp, i, i, vp, i, i, vp, i, i, vp, i, vp]
    L.glue_Eutropia_diffChangePar.argtypes = [vp, i, i, vp, i, i, vp, i, i, vp, i, vp, i, i, vp]
    L.glue_session_round.argtypes = [vp, i, vp, i, vp, i, vp, vp, i, vp, vp]
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def call_par(L, mat_in, mat_out, conc, niter, indvar, adj_cols=2):
    adj = np.require(mat_in, dtype=np.int32, requirements=["F", "A"])
    nn = np.require(mat_out, dtype=np.int32, requirements=["F", "A"])
    c = np.require(conc, dtype=np.float64, requirements=["F", "A"])
    ni, iv = np.ascontiguousarray(niter, np.int32), np.ascontiguousarray(indvar, np.int32)
    out = np.full(c.shape, np.nan, order="F")
    rc = L.glue_Eutropia_diffChangePar(_p(adj), adj.shape[0], adj_cols, _p(nn), nn.shape[0], 2, _p(c), c.shape[0], c.shape[1],
                                       _p(ni), ni.size, _p(iv), iv.size, 1, _p(out))
    return rc, out, L.glue_last_error().decode()


def test_glue_builds_loads_and_raises_r_level_errors(glue):
    for name in ("glue_Eutropia_diffChange", "glue_Eutropia_diffChangePar", "glue_session_round", "glue_last_error"):
        assert hasattr(glue, name)
    rng = np.random.default_rng(0)
    mat_in = np.array([[1, 2], [2, 1]], np.int32)
    mat_out = np.array([[1, 2], [2, 2]], np.int32)
    conc = rng.uniform(0, 1, (2, 2))
    rc, _, msg = call_par(glue, np.zeros((2, 3), np.int32), mat_out, conc, [1, 1], [1, 2], adj_cols=3)
    assert rc == 1 and "two columns" in msg                           # Rcpp::stop -> R error
    rc, _, msg = call_par(glue, mat_in, mat_out, conc, [1], [1, 2])
    assert rc == 1 and "index out of bounds" in msg                   # niter(i) past the end, src/diffChange.cpp:48
    import eutropia_b200 as eb
    if eb._lib.lib().eut_device_count() == 0:
        rc, _, msg = call_par(glue, mat_in, mat_out, conc, [1, 1], [1, 2])
        assert rc == 1 and "no CUDA device" in msg                    # no CPU path behind the glue either


@pytest.mark.gpu
def test_glue_reference_golden_vectors(glue):
    g = np.load(os.path.join(GOLD, "diffusion_reference.npz"))
    rc, out, msg = call_par(glue, g["a_mat_in"], g["a_mat_out"], g["a_conc"], g["a_niter"], g["a_indvar"])
    assert rc == 0, msg
    assert np.array_equal(out, g["a_expected"])
    os.environ["EUTROPIA_B200_DEVICES"] = "0"
    try:
        rc, out, msg = call_par(glue, g["a_mat_in"], g["a_mat_out"], g["a_conc"], g["a_niter"], g["a_indvar"])
    finally:
        del os.environ["EUTROPIA_B200_DEVICES"]
    assert rc == 0 and np.array_equal(out, g["a_expected"])
    for k in ("b", "c"):                                               # serial entry point, doubles throughout
        adj = np.require(g[k + "_mat_in"], dtype=np.float64, requirements=["F", "A"])
        nn = np.require(g[k + "_mat_out"], dtype=np.float64, requirements=["F", "A"])
        conc = np.require(g[k + "_conc"], dtype=np.float64, requirements=["F", "A"])
        nit = np.ascontiguousarray(g[k + "_niter"], np.float64)
        out = np.full(conc.shape, np.nan, order="F")
        rc = glue.glue_Eutropia_diffChange(_p(adj), adj.shape[0], 2, _p(nn), nn.shape[0], 2, _p(conc), conc.shape[0], conc.shape[1],
                                           _p(nit), nit.size, _p(out))
        assert rc == 0, glue.glue_last_error().decode()
        want = g[k + "_expected"]
        assert np.array_equal(np.isnan(out), np.isnan(want)) and np.array_equal(out[~np.isnan(want)], want[~np.isnan(want)])
    # an index outside 1..N: Armadillo's bounds check in the reference, an R error here as well
    bad = np.array(g["a_mat_in"], np.int32); bad[7, 0] = g["a_mat_out"].shape[0] + 5
    rc, _, msg = call_par(glue, bad, g["a_mat_out"], g["a_conc"], g["a_niter"], g["a_indvar"])
    assert rc == 1 and "out of bounds" in msg


@pytest.mark.gpu
def test_glue_session_round(glue, small_mesh):
    m = small_mesh
    rng = np.random.default_rng(4)
    conc = np.asfortranarray(rng.uniform(0, 25, (m.nfields, 4)))
    niter, indvar = np.array([9, 3, 9], np.int32), np.array([1, 2, 4], np.int32)
    mat_in = np.require(m.mat_in, dtype=np.int32, requirements=["F", "A"])
    mat_out = np.require(m.mat_out, dtype=np.int32, requirements=["F", "A"])
    out = np.full(conc.shape, np.nan, order="F")
    col_max = np.zeros(4)
    rc = glue.glue_session_round(_p(mat_in), mat_in.shape[0], _p(mat_out), m.nfields, _p(conc), 4, _p(niter), _p(indvar), 3,
                                 _p(out), _p(col_max))
    assert rc == 0, glue.glue_last_error().decode()
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 2)
    assert np.array_equal(out, ref) and np.array_equal(col_max, ref.max(axis=0))
