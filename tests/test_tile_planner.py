"""CPU suite: the tile planner (eutropia_b200/csrc/tiles.cu) through a host-only plan.  The tables the
tiled kernel consumes are replayed in numpy: every tile's shared-memory image is assembled from its
copy lists and every neighbour slot must then read exactly the value mat.in prescribes, in mat.in
order."""
import numpy as np
import pytest

import eutropia_b200 as eb
from eutropia_b200 import _lib, mesh


def replay(m, plan):
    inf = plan.info
    assert inf.tiled == 1
    n = m.nfields
    perm = plan.perm()
    assert np.array_equal(np.sort(perm), np.arange(n))            # a permutation
    meta, copies, loc = plan.tiles()
    assert meta[0, 0] == 0 and np.array_equal(meta[1:, 0], np.cumsum(meta[:-1, 1]))
    assert meta[:, 1].sum() == n and meta[:, 1].max() <= 1536 and meta[:, 1].min() > 0
    assert meta[:, 3].max() <= 64 and meta[:, 5].max() <= 512
    # reference CSR (stable by `to`)
    to = m.mat_in[:, 1].astype(np.int64) - 1
    order = np.argsort(to, kind="stable")
    frm = (m.mat_in[:, 0] - 1)[order]
    row_ptr = np.concatenate([[0], np.cumsum(np.bincount(to, minlength=n))])
    iperm = np.empty(n, dtype=np.int64); iperm[perm] = np.arange(n)
    value = np.random.default_rng(0).uniform(1, 2, n)              # value of internal position q
    vq = value[iperm]
    for t in range(meta.shape[0]):
        t0, cnt, b0, nb, s0, ns, tx, img = meta[t]
        image = np.full(img, np.nan)
        image[0:2] = 0.0
        moved = 0
        for g, y in copies[b0:b0 + nb]:
            so, npts = y & 0xffff, y >> 16
            assert g % 2 == 0 and so % 2 == 0 and npts % 2 == 0 and npts > 0   # 16-byte rule of cp.async.bulk
            assert np.all(np.isnan(image[so:so + npts]))
            image[so:so + npts] = vq[g:g + npts]
            moved += 8 * npts
        assert moved == tx
        for g, so in copies[s0:s0 + ns]:
            assert np.isnan(image[so])
            image[so] = vq[g]
        own_off = 2 + (t0 & 1)
        assert np.array_equal(image[own_off:own_off + cnt], vq[t0:t0 + cnt])
        for lp in range(0, cnt, max(1, cnt // 97)):                # sample of the tile's points
            q = t0 + lp
            i = iperm[q]
            nb_ref = frm[row_ptr[i]:row_ptr[i + 1]]
            offs = loc[:, q].astype(np.int64)
            assert np.all(offs % 8 == 0)
            got = image[offs // 8]
            assert np.array_equal(got[:nb_ref.size], value[nb_ref])        # mat.in order, right values
            assert np.all(got[nb_ref.size:] == 0.0)                        # unused slots -> zero cell
    return inf


@pytest.mark.parametrize("name,layers", [("Rectangle_40_33", 3), ("Petri_23", 6), ("harbor", 4), ("Rectangle_70_12", 1)])
def test_tile_tables_replay(name, layers):
    poly = mesh.harbor_polygon(90.0) if name == "harbor" else name
    m = mesh.make_mesh(poly, 1.0, layers)
    plan = eb.Plan(m.mat_in, m.mat_out, order=_lib.ORDER_TILED, host_only=True)
    inf = replay(m, plan)
    assert inf.halo_ratio < 3.0


def test_tiles_fill_up_on_large_mesh():
    m = mesh.make_mesh("Rectangle_300_300", 1.0, 3)
    plan = eb.Plan(m.mat_in, m.mat_out, order=_lib.ORDER_TILED, host_only=True)
    inf = plan.info
    meta, _, _ = plan.tiles()
    assert inf.tiled == 1 and np.median(meta[:, 1]) > 0.8 * inf.max_tile_points   # tiles are close to capacity
    assert inf.halo_ratio < 0.7


def test_non_lattice_graph_falls_back():
    rng = np.random.default_rng(2)
    n, e = 3000, 20000
    mat_in = np.unique(np.stack([rng.integers(1, n + 1, e), rng.integers(1, n + 1, e)], axis=1), axis=0).astype(np.int32)
    deg_in = np.bincount(mat_in[:, 1], minlength=n + 1)
    mat_in = mat_in[deg_in[mat_in[:, 1]] <= 12]
    mat_out = np.stack([np.arange(1, n + 1), np.bincount(mat_in[:, 0], minlength=n + 1)[1:] + 1], axis=1).astype(np.int32)
    plan = eb.Plan(mat_in, mat_out, n_points=n, order=_lib.ORDER_TILED, host_only=True)
    inf = plan.info
    # random graph: either it still tiles within the limits or the planner declines and the ELL path is used
    assert inf.tiled in (0, 1) and inf.order in (_lib.ORDER_TILED, _lib.ORDER_ROWBLOCK)
    assert np.array_equal(np.sort(plan.perm()), np.arange(n))
