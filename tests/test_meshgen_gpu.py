"""GPU parity test of the device neighbour-map construction (eut_build_dcm, csrc/meshgen.cu) against
the host mirror of `build.DCM` (mesh.build_dcm, itself pinned by tests/test_oracle.py): integer index
maps, so the bar is bit-exact equality -- including the snapping quirks of `correct_numbers` with a
field size that is not exactly representable."""
import time

import numpy as np
import pytest

from eutropia_b200 import mesh

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,layers,fs", [("Rectangle_40_33", 3, 1.0), ("Petri_23", 6, 1.0), ("harbor", 4, 1.0),
                                            ("Rectangle_70_12", 1, 1.0), ("Petri_30", 3, 0.7), ("Rectangle_25_20", 2, 1.3)])
def test_build_dcm_bit_exact(name, layers, fs):
    poly = mesh.harbor_polygon(90.0) if name == "harbor" else mesh.universe_polygon_preset(name)
    pts = mesh.make_field_points(poly, fs, layers)
    mat_in, mat_out = mesh.build_dcm(pts, fs)
    g_in, g_out = mesh.build_dcm_gpu(pts, fs)
    assert g_in.dtype == np.int32 and g_in.shape == mat_in.shape
    assert np.array_equal(g_in, mat_in) and np.array_equal(g_out, mat_out)


def test_build_dcm_at_scale_properties():
    poly = mesh.harbor_polygon(420.0)
    pts = mesh.make_field_points(poly, 1.0, 3)
    t = time.perf_counter()
    g_in, g_out = mesh.build_dcm_gpu(pts, 1.0)
    dt = time.perf_counter() - t
    n = pts.shape[0]
    assert n > 300_000 and dt < 20.0
    frm, to = g_in[:, 0].astype(np.int64), g_in[:, 1].astype(np.int64)
    assert frm.min() >= 1 and to.max() <= n
    key = to * (n + 1) + frm
    assert np.all(np.diff(key) > 0)                                   # ordered by (to, from), no duplicates
    rev = np.sort(frm * (n + 1) + to)
    assert np.array_equal(rev, key)                                   # the lattice is symmetric
    deg = np.bincount(frm, minlength=n + 1)[1:]
    assert deg.max() == 12 and np.array_equal(g_out[:, 1], deg + 1) and np.array_equal(g_out[:, 0], np.arange(1, n + 1))
    d = pts[to - 1] - pts[frm - 1]
    assert np.allclose(np.abs(d[:, 2]) + np.abs(d[:, 0]) + np.abs(d[:, 1]), np.where(d[:, 2] == 0, 1.0, 2.0))
