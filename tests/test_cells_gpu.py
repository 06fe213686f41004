"""GPU parity tests of the cell<->field exchange (cell map, claim rescale, gather, scatter+clamp)
through the C ABI against the CPU oracle.  Index maps (field ids per cell, in key order) must be
bit-exact; values must agree within 1e-12 relative (sums that R takes in long double are taken as
hi+lo pairs on the device), and the scatter must be reproducible bit for bit run to run."""
import os

import numpy as np
import pytest

import oracle
import eutropia_b200 as eb
from eutropia_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
REL = 1e-12


def close(a, b, rel=REL, absol=0.0):
    return np.all(np.abs(a - b) <= rel * np.maximum(np.abs(b), 1e-300) + absol)


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "cells_small.npz"))


def test_cellmap_bit_exact_indices(gold):
    g = gold
    plan = eb.Plan(g["mat_in"], g["mat_out"])
    cells = eb.Cells(plan)
    cells.build_lists(g["field_pts"], g["cx"], g["cy"], g["size"], g["scv"])
    ptr, fid, fd, acc = cells.get_lists()
    assert np.array_equal(ptr, g["cell_ptr"])
    assert np.array_equal(fid, g["field_id"])            # index map: bit-exact, key order (z,x,y)
    assert np.array_equal(fd, g["field_dist"])           # IEEE sqrt / mul / add: bit-exact too
    assert np.array_equal(acc, g["acc_raw"])
    cells.claim_rescale()
    _, _, _, acc2 = cells.get_lists()
    assert close(acc2, g["acc_prop"])
    assert np.array_equal(acc2[g["acc_prop"] == g["acc_raw"]], g["acc_raw"][g["acc_prop"] == g["acc_raw"]])


def test_cellmap_xyz_mode_and_no_cells(gold):
    g = gold
    plan = eb.Plan(g["mat_in"], g["mat_out"])
    cells = eb.Cells(plan)
    cells.build_lists(g["field_pts"], g["cx"], g["cy"], g["size"], g["scv"], dist_mode=1)
    ptr, fid, fd, acc = cells.get_lists()
    rp, rf, rd, ra = oracle.cellmap(g["field_pts"], g["cx"], g["cy"], g["size"], g["scv"], dist_mode=1)
    assert np.array_equal(ptr, rp) and np.array_equal(fid, rf) and np.array_equal(fd, rd) and np.array_equal(acc, ra)
    cells.build_lists(g["field_pts"], [], [], [], [])
    assert cells.n_entries == 0
    # a cell far outside the mesh has an empty list
    cells.build_lists(g["field_pts"], [1e6, float(g["cx"][0])], [1e6, float(g["cy"][0])], [1.2, 1.2], [3.1, 3.1])
    ptr, fid, _, _ = cells.get_lists()
    assert ptr[1] == 0 and ptr[2] == fid.size > 0


def test_gather_scatter_match_oracle(gold):
    g = gold
    vol = float(g["field_vol"])
    plan = eb.Plan(g["mat_in"], g["mat_out"])
    field = eb.Field(plan, g["conc"].shape[1])
    field.upload(g["conc"])
    cells = eb.Cells(plan)
    cells.set_lists(g["cell_ptr"], g["field_id"], g["field_dist"], g["acc_prop"])
    fmol = cells.gather(field, vol)
    assert close(fmol, g["fmol"])
    cells.scatter(field, vol, g["is_const"], g["ex_ptr"], g["ex_cpd"], g["ex_flx"])
    after = field.download()
    ref = g["after"]
    # tolerance relative to the field magnitude: uptake and production cancel inside one sum
    assert close(after, ref, rel=REL, absol=REL * np.abs(g["conc"]).max())
    assert np.array_equal(after == 0.0, ref == 0.0)                 # the < 1e-12 clamp hits the same entries
    assert np.array_equal(after[:, 4], g["conc"][:, 4])             # constant compound untouched
    untouched = np.ones(g["conc"].shape[0], bool); untouched[g["field_id"] - 1] = False
    assert np.array_equal(after[untouched], g["conc"][untouched])
    # run-to-run reproducibility (deterministic order, no atomics)
    field.upload(g["conc"])
    cells.scatter(field, vol, g["is_const"], g["ex_ptr"], g["ex_cpd"], g["ex_flx"])
    assert np.array_equal(field.download(), after)


def test_round_scatter_diffuse_gather(gold):
    # one simulated round on the resident field == the oracle's round
    g = gold
    vol = float(g["field_vol"])
    plan = eb.Plan(g["mat_in"], g["mat_out"])
    field = eb.Field(plan, 6)
    field.upload(g["conc"])
    cells = eb.Cells(plan)
    cells.build_lists(g["field_pts"], g["cx"], g["cy"], g["size"], g["scv"])
    cells.claim_rescale()
    cells.scatter(field, vol, g["is_const"], g["ex_ptr"], g["ex_cpd"], g["ex_flx"])
    field.diffuse([60, 22, 60, 60], [1, 2, 3, 6])
    fmol = cells.gather(field, vol)
    ref = oracle.diffChangePar(g["mat_in"], g["mat_out"], g["after"], [60, 22, 60, 60], [1, 2, 3, 6], 2)
    rfmol = oracle.gather(ref, g["cell_ptr"], g["field_id"], g["acc_prop"], vol)
    scale = np.abs(g["conc"]).max()
    assert close(field.download(), ref, absol=REL * scale)
    assert close(fmol, rfmol, rel=1e-11, absol=1e-12 * np.abs(rfmol).max())


def test_scatter_nan_and_clamp_rules(small_mesh):
    m = small_mesh
    plan = eb.Plan(m.mat_in, m.mat_out)
    conc = np.full((m.nfields, 3), 2.0)
    conc[:, 1] = 0.0                        # nothing accessible: colSum 0 -> 0/0 -> NaN -> dropped
    cx, cy = np.array([0.0, 0.4]), np.array([0.0, 0.2])
    size, scv = np.full(2, 1.24), np.full(2, 3.1)
    cells = eb.Cells(plan)
    cells.build_lists(m.field_pts, cx, cy, size, scv)
    ptr, fid, fd, acc = cells.get_lists()
    ex_ptr = np.array([0, 3, 4]); ex_cpd = np.array([2, 1, 3, 2]); ex_flx = np.array([-1e-3, -1e-3, 2e-3, 5e-3])
    ref = oracle.scatter(conc, None, ptr, fid, fd, acc, m.field_vol, ex_ptr, ex_cpd, ex_flx)
    field = eb.Field(plan, 3)
    field.upload(conc)
    cells.scatter(field, m.field_vol, None, ex_ptr, ex_cpd, ex_flx)
    got = field.download()
    assert close(got, ref, absol=1e-12 * 2.0)
    assert not np.isnan(got).any()
    # fields shared by both cells lose cell 1's production of compound 2 as well (NaN swallows the group sum)
    both = np.intersect1d(fid[ptr[0]:ptr[1]], fid[ptr[1]:ptr[2]]) - 1
    assert both.size > 0 and np.all(got[both, 1] == 0.0)


def test_properties_at_scale():
    m = synth.config_mesh("harbor:200000")
    n_cells, C = 5000, 16
    conc = synth.synthetic_concentrations(m.field_pts, C, seed=51)
    cx, cy, size, scv = synth.place_cells(m, n_cells, seed=52)
    plan = eb.Plan(m.mat_in, m.mat_out)
    field = eb.Field(plan, C)
    field.upload(conc)
    cells = eb.Cells(plan)
    cells.build_lists(m.field_pts, cx, cy, size, scv)
    ptr, fid, fd, acc = cells.get_lists()
    # sample of cells against the oracle (index map exact)
    pick = np.array([0, 1, 77, 4999])
    rp, rf, rd, ra = oracle.cellmap(m.field_pts, cx[pick], cy[pick], size[pick], scv[pick])
    for j, i in enumerate(pick):
        assert np.array_equal(fid[ptr[i]:ptr[i + 1]], rf[rp[j]:rp[j + 1]])
        assert np.array_equal(fd[ptr[i]:ptr[i + 1]], rd[rp[j]:rp[j + 1]])
    per_cell = np.diff(ptr)
    assert 100 < per_cell.mean() < 160                      # ~133 fields per cell at 3 layers
    cells.claim_rescale()
    _, _, _, acc2 = cells.get_lists()
    claim = np.bincount(fid, weights=acc2, minlength=m.nfields + 1)
    assert claim.max() <= 1 + 1e-9
    assert close(acc2, oracle.claim(fid, acc, m.nfields), rel=1e-11)
    # gather vs oracle at full size
    fmol = cells.gather(field, m.field_vol)
    assert close(fmol, oracle.gather(conc, ptr, fid, acc2, m.field_vol))
    # scatter: total fmol moved per compound equals the sum of the cells' fluxes (small fluxes: no clamp)
    ex_ptr, ex_cpd, ex_flx = synth.synthetic_exchanges(n_cells, C, 4, seed=53)
    ex_flx *= 1e-4
    cells.scatter(field, m.field_vol, None, ex_ptr, ex_cpd, ex_flx)
    after = field.download()
    moved = (after - conc).sum(0) * 1e12 * (m.field_vol / 1e15)
    want = np.bincount(ex_cpd - 1, weights=ex_flx, minlength=C)
    assert np.allclose(moved, want, rtol=1e-6, atol=1e-9)
    ref = oracle.scatter(conc, None, ptr, fid, fd, acc2, m.field_vol, ex_ptr, ex_cpd, ex_flx)
    assert close(after, ref, absol=1e-12 * np.abs(conc).max())


def test_vignette_like_rounds():
    """BASELINE.json configs[0] in shape: the Crossfeeding vignette's environment (`Petri_80`, field
    size 1, 6 layers, vignettes/Crossfeeding.Rmd:31-33: 124 k points; 171 compound columns of which
    19 are constant medium components, inst/extdata/medium.csv:2-20; 30 cells; D = 75, deltaTime
    1/6 h -> 60 substeps), three rounds of  cell map -> claim -> gather -> (synthetic fluxes instead
    of the FBA) -> scatter + clamp -> diffuse_compoundsPar  on the resident field, against the same
    rounds composed from the oracle."""
    from eutropia_b200 import environment, mesh as mesh_mod
    m = mesh_mod.make_mesh("Petri_80", 1.0, 6)
    assert 100_000 < m.nfields < 150_000
    n_cols, n_const, n_cells = 171, 19, 30
    rng = np.random.default_rng(80)
    conc = np.zeros((m.nfields, n_cols))
    conc[:, :n_const] = 1000.0                                   # constant medium components
    conc[:, n_const:n_const + 7] = rng.uniform(0.5, 25.0, 7)     # the non-constant medium, uniform at t = 0
    is_const = np.zeros(n_cols, dtype=bool); is_const[:n_const] = True
    D = np.full(n_cols, 75.0)
    env = environment.GrowthEnvironment(m, [f"cpd{i}" for i in range(n_cols)], conc, D, is_const)
    vol, dt = m.field_vol, 1 / 6
    ref = conc.copy()
    cx, cy, size, scv = synth.place_cells(m, n_cells, seed=7)
    for rnd in range(3):
        cx = cx + rng.normal(0, 0.4, n_cells); cy = cy + rng.normal(0, 0.4, n_cells)      # cells drift
        ex_ptr, ex_cpd, ex_flx = synth.synthetic_exchanges(n_cells, n_cols, 12, seed=100 + rnd)
        ex_flx = ex_flx * 1e-3
        # device
        env.cells.build_lists(m.field_pts, cx, cy, size, scv)
        env.cells.claim_rescale()
        fmol = env.cells.gather(env.field, vol)
        env.cells.scatter(env.field, vol, is_const, ex_ptr, ex_cpd, ex_flx)
        env.diffuse_compoundsPar(dt, n_cores=4)
        # oracle
        ptr, fid, fd, acc = oracle.cellmap(m.field_pts, cx, cy, size, scv)
        acc = oracle.claim(fid, acc, m.nfields)
        d_ptr, d_fid, d_fd, d_acc = env.cells.get_lists()
        assert np.array_equal(d_ptr, ptr) and np.array_equal(d_fid, fid)                     # index maps: bit-exact
        rfmol = oracle.gather(ref, ptr, fid, acc, vol)
        assert close(fmol, rfmol, rel=1e-11, absol=1e-12 * max(np.abs(rfmol).max(), 1e-300))
        ref = oracle.scatter(ref, is_const, ptr, fid, fd, acc, vol, ex_ptr, ex_cpd, ex_flx)
        sd_pos = (ref.max(0) > ref.min(0)) & ~is_const
        ind = np.flatnonzero(sd_pos) + 1
        niter = environment.diffusion_niter(D[ind - 1], dt, m.field_size)
        assert np.all(niter == 60)
        ref = oracle.diffChangePar(m.mat_in, m.mat_out, ref, niter.astype(np.int32), ind.astype(np.int32), 8)
        got = env.concentrations
        assert np.array_equal(got[:, :n_const], conc[:, :n_const])                           # constants never move
        assert close(got, ref, absol=REL * 25.0), float(np.abs(got - ref).max())
    assert ind.size > 12                                          # the diffusing set grows with the exchanges


def test_chemotaxis_shift_matches_oracle(small_mesh):
    """Chemotaxis vector of every cell (R/run_simulation.R:260-291) from the device moments, against the
    statement-by-statement numpy restatement; includes cells whose neighbourhood holds none of the
    compound (uniform weights, zero sensing)."""
    m = small_mesh
    rng = np.random.default_rng(12)
    conc = rng.uniform(0, 5, (m.nfields, 3))
    conc[m.field_pts[:, 0] < 0, 1] = 0.0                    # the left half holds nothing of compound 2
    plan = eb.Plan(m.mat_in, m.mat_out)
    field = eb.Field(plan, 3)
    field.upload(conc)
    cx, cy, size, scv = synth.place_cells(m, 40, seed=3)
    cells = eb.Cells(plan)
    cells.build_lists(m.field_pts, cx, cy, size, scv)
    ptr, fid, _, _ = cells.get_lists()
    from eutropia_b200 import exchange
    got = exchange.chemotaxis_shift(cells, field, m.field_pts, 2, cx, cy, size, float(scv[0]), 0.1, 1.2, 0.01, 11.0, 1 / 6)
    ref = oracle.chemotaxis_shift(conc[:, 1], m.field_pts, ptr, fid, cx, cy, size, float(scv[0]), 0.1, 1.2, 0.01, 11.0, 1 / 6)
    scale = max(np.abs(ref[0]).max(), np.abs(ref[1]).max(), 1e-300)
    assert np.all(np.abs(got[0] - ref[0]) <= REL * scale) and np.all(np.abs(got[1] - ref[1]) <= REL * scale)
    assert np.any(ref[0] != 0) and np.any(ref[0] == 0)       # both branches are exercised


@pytest.mark.parametrize("n_cols", [5, 30, 70])
def test_scatter_generations_agree(harbor_mesh, n_cols, monkeypatch):
    """The point-per-lane scatter (default) against the two earlier generations: the same terms in the same
    cell order per (field, compound) sum, so the fields must be equal bit for bit; 70 columns take two
    passes (64 + 6).  With the colSums taken over from a gather of the same field state (instead of a
    second pass over the field) the result stays inside the 1e-12 bar."""
    m = harbor_mesh
    rng = np.random.default_rng(60 + n_cols)
    conc = synth.synthetic_concentrations(m.field_pts, n_cols, seed=61)
    conc[rng.uniform(size=conc.shape) < 0.02] = 0.0
    n_cells = 400
    cx, cy, size, scv = synth.place_cells(m, n_cells, seed=62)
    ex_ptr, ex_cpd, ex_flx = synth.synthetic_exchanges(n_cells, n_cols, min(9, n_cols), seed=63)
    ex_flx *= 1e-3
    ex_flx[::17] = 0.0                                   # exchanges without a contribution
    is_const = np.zeros(n_cols, bool); is_const[2] = True
    plan = eb.Plan(m.mat_in, m.mat_out)
    cells = eb.Cells(plan)
    cells.build_lists(m.field_pts, cx, cy, size, scv)
    cells.claim_rescale()
    out = {}
    for gen in ("1", "2", "3"):
        monkeypatch.setenv("EUT_SCATTER_GEN", gen)
        field = eb.Field(plan, n_cols)
        field.upload(conc)
        cells.scatter(field, m.field_vol, is_const, ex_ptr, ex_cpd, ex_flx)
        out[gen] = field.download()
        field.close()
    assert np.array_equal(out["3"], out["2"]) and np.array_equal(out["3"], out["1"])
    assert np.any(out["3"] != conc)
    monkeypatch.delenv("EUT_SCATTER_GEN")
    field = eb.Field(plan, n_cols)
    field.upload(conc)
    cells.gather(field, m.field_vol)                     # colSums of the next scatter
    cells.scatter(field, m.field_vol, is_const, ex_ptr, ex_cpd, ex_flx)
    again = field.download()
    assert close(again, out["3"], absol=1e-12 * np.abs(conc).max())
    assert np.array_equal(again == 0.0, out["3"] == 0.0)
    # a field that changed after the gather must not use its sums
    field.upload(conc * 2.0)
    cells.scatter(field, m.field_vol, is_const, ex_ptr, ex_cpd, ex_flx)
    ptr, fid, fd, acc = cells.get_lists()
    ref = oracle.scatter(conc * 2.0, is_const, ptr, fid, fd, acc, m.field_vol, ex_ptr, ex_cpd, ex_flx)
    assert close(field.download(), ref, absol=1e-12 * np.abs(conc).max() * 2)
