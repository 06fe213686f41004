"""CPU suite: the oracle against (1) outputs of the reference itself -- src/diffChange.cpp compiled
unmodified into oracle/_ref, and the vectors it produced in the build container
(tests/golden/diffusion_reference.npz, tests/golden/make_reference_golden.py); (2) hand-derived
known answers and the invariants the reference implies but never tests.  The
reference ships no golden vectors of its own; the fixtures made by tests/golden/make_golden.py come
from the oracle and only guard the R-language steps (cell lists, gather, scatter) against
regressions -- parity is unpinned for those."""
import os

import numpy as np
import pytest

import oracle
from eutropia_b200 import mesh

B = 1.0 / 13


def tiny_graph():
    # 3 points on a line: 1 <-> 2 <-> 3 ; mat.in ordered by (to, from)
    mat_in = np.array([[2, 1], [1, 2], [3, 2], [2, 3]], dtype=np.int32)
    mat_out = np.array([[1, 2], [2, 3], [3, 2]], dtype=np.int32)  # out-degree + 1
    return mat_in, mat_out


def test_known_answer_one_substep():
    mat_in, mat_out = tiny_graph()
    c = np.array([[13.0], [26.0], [39.0]])
    out = oracle.diffChangePar(mat_in, mat_out, c, [1], [1], 1)
    # hand evaluation in the reference's operation order (src/diffChange.cpp:50-63)
    y1 = (0.0 + 26.0) * B - 13.0 * ((2.0 - 1) * B)
    y2 = ((0.0 + 13.0) + 39.0) * B - 26.0 * ((3.0 - 1) * B)
    y3 = (0.0 + 26.0) * B - 39.0 * ((2.0 - 1) * B)
    exp = np.array([[13.0 + y1], [26.0 + y2], [39.0 + y3]])
    assert np.array_equal(out, exp)
    assert out[:, 0].tolist() == [14.0, 26.0, 38.0]


def test_serial_equals_par_and_double_indices(small_mesh):
    m = small_mesh
    rng = np.random.default_rng(3)
    conc = rng.uniform(0, 25, (m.nfields, 5))
    niter = [4, 0, 7, 1, 3]
    a = oracle.diffChange(m.mat_in.astype(np.float64), m.mat_out.astype(np.float64), conc, niter)
    b = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, [1, 2, 3, 4, 5], 3)
    assert np.array_equal(a, b)
    assert np.array_equal(a[:, 1], conc[:, 1])          # niter 0: untouched


def test_niter_shorter_than_columns_and_fractional(small_mesh):
    m = small_mesh
    conc = np.random.default_rng(4).uniform(0, 1, (m.nfields, 3))
    a = oracle.diffChange(m.mat_in, m.mat_out, conc, [2.5])       # k < 2.5 -> 3 substeps
    b = oracle.diffChangePar(m.mat_in, m.mat_out, conc, [3], [1], 1)
    assert np.array_equal(a, b)
    assert np.array_equal(a[:, 1:], conc[:, 1:])


def test_mass_conservation_and_fixed_point(harbor_mesh):
    m = harbor_mesh
    rng = np.random.default_rng(5)
    conc = rng.uniform(0, 25, (m.nfields, 2))
    out = oracle.diffChangePar(m.mat_in, m.mat_out, conc, [25, 25], [1, 2], 2)
    assert np.allclose(out.sum(0), conc.sum(0), rtol=1e-12)
    assert out.min() >= 0
    uni = np.full((m.nfields, 1), 7.25)
    out = oracle.diffChangePar(m.mat_in, m.mat_out, uni, [10], [1], 1)
    assert np.allclose(out, uni, rtol=1e-14)


def test_out_of_range_index_raises():
    mat_in, mat_out = tiny_graph()
    bad = mat_in.copy(); bad[0, 0] = 4
    with pytest.raises(oracle.OracleIndexError):
        oracle.diffChangePar(bad, mat_out, np.ones((3, 1)), [1], [1], 1)
    with pytest.raises(oracle.OracleIndexError):
        oracle.diffChangePar(mat_in, mat_out, np.ones((3, 1)), [1], [2], 1)
    with pytest.raises(oracle.OracleIndexError):
        oracle.diffChange(mat_in, mat_out, np.ones((3, 1)), [1, 1])
    # a bad column that never runs a substep is never dereferenced
    out = oracle.diffChangePar(mat_in, mat_out, np.ones((3, 1)), [0], [9], 1)
    assert np.array_equal(out, np.ones((3, 1)))


def test_mesh_matches_reference_rules():
    m = mesh.make_mesh("Rectangle_181_181", 1.0, 3)
    assert m.nfields == 100467                    # 183 x 183 x 3
    deg_in = np.bincount(m.mat_in[:, 1], minlength=m.nfields + 1)[1:]
    assert deg_in.max() == 12 and (m.mat_out[:, 1] - 1).max() == 12
    # ordered by (to, from); mat.out ids ascending
    key = m.mat_in[:, 1].astype(np.int64) * (m.nfields + 1) + m.mat_in[:, 0]
    assert np.all(np.diff(key) > 0)
    assert np.array_equal(m.mat_out[:, 0], np.arange(1, m.nfields + 1))
    # interior point of the middle layer: 4 in-plane + 4 upper + 4 lower
    n_base = m.nfields // 3
    p = n_base + 183 * 90 + 90 + 1
    nb = m.mat_in[m.mat_in[:, 1] == p, 0]
    assert nb.size == 12
    d = m.field_pts[nb - 1] - m.field_pts[p - 1]
    assert np.allclose(np.sort(np.linalg.norm(d[:, :2], axis=1)), [np.sqrt(0.5)] * 8 + [1.0] * 4)
    assert abs(m.field_vol - 0.7071067811865476) < 1e-15   # fieldVol at fs = 1


def test_cellmap_claim_gather_scatter_small(small_mesh):
    m = small_mesh
    pts = m.field_pts
    cx = np.array([0.3, 0.9, -7.0]); cy = np.array([0.1, 0.4, 5.0])
    size = np.full(3, 1.24); scv = np.full(3, 3.1)
    ptr, fid, fd, acc = oracle.cellmap(pts, cx, cy, size, scv)
    assert ptr[0] == 0 and ptr[-1] == fid.size and fid.size > 100
    R = size[0] / 2 + scv[0]
    for i in range(3):
        sl = slice(ptr[i], ptr[i + 1])
        d = np.hypot(pts[fid[sl] - 1, 0] - cx[i], pts[fid[sl] - 1, 1] - cy[i])
        assert np.allclose(d, fd[sl]) and d.max() <= R
        # key order (z, x, y)
        k = pts[fid[sl] - 1]
        order = np.lexsort((k[:, 1], k[:, 0], k[:, 2]))
        assert np.array_equal(order, np.arange(order.size))
        # brute force membership
        dall = np.hypot(pts[:, 0] - cx[i], pts[:, 1] - cy[i])
        assert set(fid[sl].tolist()) == set((np.flatnonzero((dall <= R) & (pts[:, 2] <= R)) + 1).tolist())
    assert acc.max() <= 1.0 and acc.min() >= 0.0
    acc2 = oracle.claim(fid, acc, m.nfields)
    claim = np.bincount(fid, weights=acc2, minlength=m.nfields + 1)
    assert claim.max() <= 1 + 1e-12                    # oversubscribed fields are scaled back to <= 1
    conc = np.random.default_rng(1).uniform(0, 25, (m.nfields, 4))
    fmol, fpf = oracle.gather(conc, ptr, fid, acc2, m.field_vol, want_fpf=True)
    assert np.allclose(fmol[0, 2], np.sum(conc[fid[ptr[0]:ptr[1]] - 1, 2] / 1000 * m.field_vol * acc2[ptr[0]:ptr[1]]))
    # scatter: uptake of cpd 1 by cell 0 removes exactly |ex| fmol (|ex| below the accessible amount, as the LP bound guarantees), production adds ex fmol
    ex_ptr = np.array([0, 2, 2, 2]); ex_cpd = np.array([1, 3]); ex_flx = np.array([-0.05, 0.25])
    out = oracle.scatter(conc, None, ptr, fid, fd, acc2, m.field_vol, ex_ptr, ex_cpd, ex_flx)
    d_fmol = (out - conc).sum(0) * 1e12 * (m.field_vol / 1e15)
    assert np.allclose(d_fmol, [-0.05, 0, 0.25, 0], atol=1e-12)
    # constant compounds are left alone; results below 1e-12 are clamped to 0
    out2 = oracle.scatter(conc, [1, 0, 0, 0], ptr, fid, fd, acc2, m.field_vol, ex_ptr, ex_cpd, ex_flx)
    assert np.array_equal(out2[:, 0], conc[:, 0]) and np.array_equal(out2[:, 2], out[:, 2])
    tiny = np.full_like(conc, 1e-13)
    out3 = oracle.scatter(tiny, None, ptr, fid, fd, acc2, m.field_vol, ex_ptr, ex_cpd, np.array([-1e-9, 1e-12]))
    touched = np.unique(fid[ptr[0]:ptr[1]]) - 1
    assert np.all(out3[touched, 0] == 0.0)


def test_golden_fixture_regression():
    path = os.path.join(os.path.dirname(__file__), "golden", "diffusion_small.npz")
    g = np.load(path)
    out = oracle.diffChangePar(g["mat_in"], g["mat_out"], g["conc"], g["niter"], g["indvar"], 2)
    assert np.array_equal(out, g["expected"])


# ---- the pin: outputs of the reference itself ------------------------------------------------------

def bits_equal(a, b):
    """Same doubles, NaN == NaN; the sign of zero is part of the contract here."""
    a, b = np.asarray(a), np.asarray(b)
    nan = np.isnan(a)
    return a.shape == b.shape and np.array_equal(nan, np.isnan(b)) and \
        np.array_equal(a[~nan].view(np.uint64), b[~nan].view(np.uint64))


def reference_golden():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "diffusion_reference.npz"))


def test_restatement_matches_reference_golden_vectors():
    g = reference_golden()
    out = oracle.diffChangePar(g["a_mat_in"], g["a_mat_out"], g["a_conc"], g["a_niter"], g["a_indvar"], 2)
    assert bits_equal(out, g["a_expected"])
    assert np.array_equal(out[:, [2, 4]], g["a_conc"][:, [2, 4]])            # not in indvar: untouched
    out = oracle.diffChange(g["b_mat_in"], g["b_mat_out"], g["b_conc"], g["b_niter"])
    assert bits_equal(out, g["b_expected"])
    # the `ychg = ychg * 0` carry-over is in the vectors: columns 3 and 4 start finite and end with NaN
    assert np.isfinite(g["b_conc"][:, 2:4]).all() and np.isnan(g["b_expected"][:, 2:4]).any()
    out = oracle.diffChange(g["c_mat_in"], g["c_mat_out"], g["c_conc"], g["c_niter"])
    assert bits_equal(out, g["c_expected"])


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref is built only where /root/reference exists")
@pytest.mark.parametrize("name,layers", [("Rectangle_30_24", 3), ("Petri_23", 6), ("harbor", 4), ("Rectangle_40_12", 1)])
def test_restatement_matches_compiled_reference(name, layers):
    m = mesh.make_mesh(mesh.harbor_polygon(40.0) if name == "harbor" else name, 1.0, layers)
    rng = np.random.default_rng(len(name) + layers)
    conc = rng.uniform(0, 30, (m.nfields, 6)) * (rng.uniform(0, 1, (m.nfields, 6)) > 0.3)
    niter = [5, 0, 17, 1, 2, 9]
    assert bits_equal(oracle.diffChange(m.mat_in, m.mat_out, conc, niter),
                      oracle.ref_diffChange(m.mat_in, m.mat_out, conc, niter))
    assert bits_equal(oracle.diffChange(m.mat_in, m.mat_out, conc, [1.5, 0.0, -3.0]),         # double bounds, short niter
                      oracle.ref_diffChange(m.mat_in, m.mat_out, conc, [1.5, 0.0, -3.0]))
    iv = [6, 1, 4, 3]
    assert bits_equal(oracle.diffChangePar(m.mat_in, m.mat_out, conc, [3, 11, 0, 6], iv, 3),
                      oracle.ref_diffChangePar(m.mat_in, m.mat_out, conc, [3, 11, 0, 6], iv, 3))
    # edge list in a different order: the summation order follows mat.in, not the point ids
    shuf = rng.permutation(m.mat_in.shape[0])
    assert bits_equal(oracle.diffChangePar(m.mat_in[shuf], m.mat_out, conc, [4], [2], 1),
                      oracle.ref_diffChangePar(m.mat_in[shuf], m.mat_out, conc, [4], [2], 1))
    # non-finite data, negative values, signed zeros
    bad = conc.copy()
    bad[3, 0] = np.inf; bad[9, 2] = np.nan; bad[:, 3] = -bad[:, 3]; bad[:, 4] = -0.0
    assert bits_equal(oracle.diffChange(m.mat_in, m.mat_out, bad, [2, 1, 2, 3, 2, 1]),
                      oracle.ref_diffChange(m.mat_in, m.mat_out, bad, [2, 1, 2, 3, 2, 1]))
    assert bits_equal(oracle.diffChangePar(m.mat_in, m.mat_out, bad, [2, 2, 2, 2, 2, 2], [1, 2, 3, 4, 5, 6], 2),
                      oracle.ref_diffChangePar(m.mat_in, m.mat_out, bad, [2, 2, 2, 2, 2, 2], [1, 2, 3, 4, 5, 6], 2))


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref is built only where /root/reference exists")
def test_compiled_reference_edge_cases():
    mat_in, mat_out = tiny_graph()
    c = np.array([[13.0], [26.0], [39.0]])
    assert oracle.ref_diffChangePar(mat_in, mat_out, c, [1], [1], 1)[:, 0].tolist() == [14.0, 26.0, 38.0]
    # empty edge list, empty niter, zero columns selected
    assert np.array_equal(oracle.ref_diffChange(np.zeros((0, 2)), mat_out, c, [3]), oracle.diffChange(np.zeros((0, 2)), mat_out, c, [3]))
    assert np.array_equal(oracle.ref_diffChange(mat_in, mat_out, c, []), c)
    assert np.array_equal(oracle.ref_diffChangePar(mat_in, mat_out, c, [], [], 1), c)
    with pytest.raises(oracle.OracleIndexError):
        oracle.ref_diffChange(np.array([[1, 4]]), mat_out, c, [1])                      # Armadillo's bounds check
    with pytest.raises(oracle.OracleIndexError):
        oracle.diffChange(np.array([[1, 4]]), mat_out, c, [1])
    with pytest.raises(oracle.OracleIndexError):
        oracle.ref_diffChange(mat_in, mat_out, c, [1, 1])                                # conc(., 1) of a 1-column matrix
    with pytest.raises(oracle.OracleIndexError):
        oracle.diffChange(mat_in, mat_out, c, [1, 1])
