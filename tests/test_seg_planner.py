"""CPU suite: the segment planner (eutropia_b200/csrc/segs.cu) through a host-only plan.  The tables
the segment kernel consumes are replayed in numpy: every tile's shared-memory image is assembled from
its copy lists (holes stay +0.0), every segment's R summation chains are evaluated in the kernel's
canonical stream order, and the result must equal -- bit for bit -- the sum over the point's mat.in
neighbours in mat.in order.  Every point must be produced exactly once."""
import numpy as np
import pytest

import eutropia_b200 as eb
from eutropia_b200 import _lib, mesh


def reference_inflow(m, value):
    """Sequential fp64 sum of value[from] per destination in mat.in row order (src/diffChange.cpp:19-21)."""
    n = m.nfields
    to = m.mat_in[:, 1].astype(np.int64) - 1
    order = np.argsort(to, kind="stable")
    frm = (m.mat_in[:, 0].astype(np.int64) - 1)[order]
    deg = np.bincount(to, minlength=n)
    row_ptr = np.concatenate([[0], np.cumsum(deg)])
    acc = np.zeros(n)
    for k in range(int(deg.max())):
        has = deg > k
        acc[has] = acc[has] + value[frm[row_ptr[:-1][has] + k]]
    return acc


def replay(m, plan, check_alignment=True, max_growth=1.6):
    inf = plan.info
    assert inf.segmented == 1
    n, R = m.nfields, inf.seg_points
    perm = plan.perm()
    npad = inf.n_padded
    assert np.unique(perm).size == n and perm.min() >= 0 and perm.max() < npad    # injective; gaps allowed
    meta, copies, segs, gens = plan.segs()
    assert meta[0, 0] >= 0 and np.all(meta[1:, 0] >= meta[:-1, 0] + meta[:-1, 1])      # disjoint, ascending
    assert n <= meta[:, 1].sum() <= npad and meta[:, 1].sum() <= max_growth * n + 64
    assert meta[:, 3].max() <= 128 and meta[:, 5].max() <= 512 and meta[:, 9].max() <= 448
    value = np.random.default_rng(0).uniform(1, 2, n)     # by reference id; sums depend on the order
    want = reference_inflow(m, value)
    iperm = np.full(npad, -1, dtype=np.int64); iperm[perm] = np.arange(n)
    vq = np.where(iperm >= 0, value[np.maximum(iperm, 0)], np.nan)   # by internal position; gap cells poison
    got = np.full(npad, np.nan)
    produced = np.zeros(npad, dtype=np.int64)
    for t in range(meta.shape[0]):
        t0, cnt, b0, nb, s0, ns, tx, img, g0, ng, x0, nx = meta[t]
        image = np.zeros(img + 1)
        filled = np.zeros(img + 1, dtype=bool)
        moved = 0
        for g, y in copies[b0:b0 + nb]:
            so, npts = y & 0xffff, y >> 16
            if check_alignment:
                assert g % 2 == 0 and so % 2 == 0 and npts % 2 == 0 and npts > 0   # 16-byte rule of cp.async.bulk
            assert not filled[so:so + npts].any()
            image[so:so + npts] = vq[g:g + npts]; filled[so:so + npts] = True
            moved += 8 * npts
        assert moved == tx
        for g, so in copies[s0:s0 + ns]:
            assert not filled[so]
            image[so] = vq[g]; filled[so] = True
        assert not filled[:inf.seg_points + 3].any()          # the zero header
        head = t0 & 1
        S = segs[g0:g0 + ng].astype(np.int64)
        if ng:
            assert np.all(S[:, list(range(7)) + list(range(8, 18))] % 8 == 0)
            ln = S[:, 7] >> 12
            out = S[:, 7] & 0xfff
            assert ln.min() >= 1 and ln.max() <= R
            C = S // 8                                           # cells
            def pair(a, col, red, j):                            # values of cells j and j+1 of a pair stream
                def cell(k):
                    inline = image[a[:, col] + k]
                    v = np.where(k == 0, image[a[:, red]], inline)
                    return np.where(k == R, image[a[:, red + 1]], v)
                return cell(j), cell(j + 1)
            for j in range(R):
                on = j < ln
                a = C[on]
                acc = np.zeros(a.shape[0])
                for col, red in ((1, 10), (2, 12)):                              # A A B B
                    v0, v1 = pair(a, col, red, j)
                    acc = acc + v0; acc = acc + v1
                acc = acc + image[a[:, 3] + j]                                       # C
                left = np.where(j == 0, image[a[:, 8]], image[a[:, 0] + j - 1])
                right = np.where(j == R - 1, image[a[:, 9]], image[a[:, 0] + j + 1])
                acc = acc + left
                acc = acc + right
                acc = acc + image[a[:, 4] + j]                                       # D
                for col, red in ((5, 14), (6, 16)):                              # E E F F
                    v0, v1 = pair(a, col, red, j)
                    acc = acc + v0; acc = acc + v1
                q = t0 + out[on] + j - head
                assert np.array_equal(image[a[:, 0] + j], vq[q])                   # own value where the kernel reads it
                got[q] = acc
                np.add.at(produced, q, 1)
        G = gens[x0:x0 + nx].astype(np.int64)
        if nx:
            acc = np.zeros(nx)
            for k in range(12):
                acc = acc + image[G[:, k]]
            q = t0 + G[:, 13] - head
            assert np.array_equal(image[G[:, 12]], vq[q])
            got[q] = acc
            np.add.at(produced, q, 1)
        assert np.array_equal(produced[t0:t0 + cnt], (iperm[t0:t0 + cnt] >= 0).astype(np.int64))
    assert np.array_equal(got[perm], want)                 # bit-exact, mat.in order
    return inf


@pytest.mark.parametrize("name,layers", [("Rectangle_40_33", 3), ("Petri_23", 6), ("harbor", 4),
                                         ("Rectangle_70_12", 1), ("Rectangle_11_10", 2), ("Petri_60", 3)])
def test_segment_tables_replay(name, layers):
    poly = mesh.harbor_polygon(90.0) if name == "harbor" else name
    m = mesh.make_mesh(poly, 1.0, layers)
    plan = eb.Plan(m.mat_in, m.mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)
    inf = replay(m, plan)
    assert inf.n_seg_covered + inf.n_generic == m.nfields
    assert inf.n_generic <= 0.08 * m.nfields + 8          # generic points only along the border (small meshes: mostly border)


def test_segments_fill_up_on_large_mesh():
    m = mesh.make_mesh("Rectangle_300_300", 1.0, 3)
    plan = eb.Plan(m.mat_in, m.mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)
    inf = plan.info
    meta, _, _, _ = plan.segs()
    assert inf.segmented == 1 and inf.n_generic <= 0.005 * m.nfields
    assert inf.n_seg_covered / (inf.n_segments * inf.seg_points) > 0.9     # segments are mostly full
    assert np.median(meta[:, 9]) > 0.8 * 224                                 # tiles are close to capacity
    assert inf.halo_ratio < 0.5


def test_shuffled_ids_still_exact():
    """A renumbered mesh is no longer raster ordered: rows are short or absent, most points become
    generic -- the tables must stay exact (or the planner declines)."""
    m = mesh.make_mesh("Rectangle_20_14", 1.0, 3)
    n = m.nfields
    rng = np.random.default_rng(4)
    relabel = np.arange(1, n + 1)
    blocks = relabel.reshape(-1, 1) if n % 7 else relabel.reshape(-1, 7)
    blocks = blocks[rng.permutation(blocks.shape[0])]
    relabel = np.empty(n, dtype=np.int64); relabel[blocks.reshape(-1) - 1] = np.arange(1, n + 1)
    mat_in = np.stack([relabel[m.mat_in[:, 0] - 1], relabel[m.mat_in[:, 1] - 1]], axis=1)
    srt = np.lexsort((mat_in[:, 0], mat_in[:, 1]))
    mat_in = mat_in[srt].astype(np.int32)
    mat_out = np.stack([np.arange(1, n + 1), np.bincount(mat_in[:, 0], minlength=n + 1)[1:] + 1], axis=1).astype(np.int32)

    class M:
        pass
    mm = M(); mm.nfields = n; mm.mat_in = mat_in; mm.mat_out = mat_out
    plan = eb.Plan(mat_in, mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)
    if plan.info.segmented:
        replay(mm, plan, check_alignment=True, max_growth=8.0)
    else:
        assert plan.info.order in (_lib.ORDER_TILED, _lib.ORDER_ROWBLOCK)


@pytest.mark.parametrize("name,layers", [("Rectangle_40_33", 3), ("Petri_23", 6), ("harbor", 4), ("Rectangle_11_10", 2)])
def test_segment_tables_decode_to_mat_in(name, layers):
    """eut_plan_get_csr on a segment plan replays the kernel's tables (here the planner's host copies; the
    -m gpu suite reads the same tables back from the device): the in-neighbour map must be mat.in, bit for bit."""
    poly = mesh.harbor_polygon(90.0) if name == "harbor" else name
    m = mesh.make_mesh(poly, 1.0, layers)
    n = m.nfields
    to, frm = m.mat_in[:, 1].astype(np.int64), m.mat_in[:, 0]
    order = np.argsort(to, kind="stable")
    row_ptr = np.concatenate([[0], np.cumsum(np.bincount(to, minlength=n + 1)[1:])])
    plan = eb.Plan(m.mat_in, m.mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)
    assert plan.info.segmented == 1
    rp, col = plan.csr()
    assert np.array_equal(rp, row_ptr) and np.array_equal(col, frm[order])


@pytest.mark.skipif(mesh.kiel_csv_path() is None, reason="the reference's inst/extdata/kiel.csv is not on this machine")
def test_kiel_polygon_of_the_reference():
    """BASELINE.json configs[2] names the reference's own non-convex polygon, `Kiel_<L>` (inst/extdata/kiel.csv
    scaled as R/growthSimulation.R:202-214).  The file is reference data and is not copied into this
    repository, so this runs where /root/reference exists (the build container): the scaling rule, the
    mesh, and the segment kernel's tables replayed bit for bit on that mesh; `Harbor_<w>` is the
    travelling stand-in of the bench."""
    raw = np.loadtxt(mesh.kiel_csv_path(), delimiter=",", skiprows=1)
    poly = mesh.universe_polygon_preset("Kiel_240")
    assert np.array_equal(poly, raw / raw[:, 0].max() * (240 / 2) / 2)           # :204-208
    assert poly[:, 0].max() == 60.0
    with pytest.raises(ValueError):
        mesh.universe_polygon_preset("Kiel_19")                                     # :210-211
    m = mesh.make_mesh("Kiel_240", 1.0, 3)
    assert 15000 < m.nfields < 30000
    plan = eb.Plan(m.mat_in, m.mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)
    inf = replay(m, plan)
    assert inf.n_seg_covered + inf.n_generic == m.nfields


def _modelled_conflicts(meta, segs):
    """Wavefronts per half-warp LDS.64, averaged over the seven streams: the largest number of lanes of a half
    warp whose cells share a residue mod 16 (same cell or the zero header = broadcast)."""
    tot = cnt = 0
    for t in range(meta.shape[0]):
        g0, ng = meta[t, 8], meta[t, 9]
        S = segs[g0:g0 + ng].astype(np.int64)
        pad = (-ng) % 16
        if pad:
            S = np.concatenate([S, np.tile(np.array([[8] + [0] * 23]), (pad, 1))])
        H = S.reshape(-1, 16, 24) // 8
        for s in range(7):
            for h in range(H.shape[0]):
                u = np.unique(H[h, :, s][H[h, :, s] > 0])
                tot += max(1, np.bincount(u % 16, minlength=16).max())
                cnt += 1
    return tot / cnt


def test_gap_rule_keeps_residues_apart(monkeypatch):
    """The coordinate gap rule (segs.cu): position - u of a piece is a function of its row index and layer, chosen
    per tile for few gap cells and a balanced histogram of segment start residues in every layer.  On a ragged
    polygon with 448-segment tiles it must beat the chain rule without the balance term on modelled bank
    conflicts, keep the gap cells below 8 % of the points, and -- the failure it was written for -- leave no
    layer of a tile with all its segments on half of the residues."""
    monkeypatch.setenv("EUT_SEG_TILE", "448")
    m = mesh.make_mesh(mesh.harbor_polygon(260.0), 1.0, 3)
    res = {}
    for name, env in (("coordinate", {"EUT_SEG_PAD": "1"}), ("chain", {"EUT_SEG_PAD": "2", "EUT_SEG_NO_BALANCE": "1"})):
        for k in ("EUT_SEG_PAD", "EUT_SEG_NO_BALANCE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        plan = eb.Plan(m.mat_in, m.mat_out, order=_lib.ORDER_SEGMENTS, host_only=True)
        inf = replay(m, plan)
        meta, _, segs, _ = plan.segs()
        res[name] = (_modelled_conflicts(meta, segs), inf.n_padded / m.nfields - 1, meta, segs)
    assert res["coordinate"][0] < res["chain"][0] - 0.1 and res["coordinate"][0] < 1.35
    assert res["coordinate"][1] < 0.08
    meta, segs = res["coordinate"][2], res["coordinate"][3]
    for t in range(meta.shape[0]):
        S = segs[meta[t, 8]:meta[t, 8] + meta[t, 9]].astype(np.int64)
        cls = (S[:, 1] > 0).astype(int) + 2 * (S[:, 5] > 0).astype(int)
        for c in np.unique(cls):
            own = (S[cls == c, 0] // 8) % 16
            if own.size >= 64:
                assert np.bincount(own, minlength=16).max() <= 2.0 * own.size / 16 + 2, (t, c, np.bincount(own, minlength=16))
