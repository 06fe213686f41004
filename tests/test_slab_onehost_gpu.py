"""eut_slab_*: one field sharded by spatial slab, driven from one process, halo values stored straight
into the neighbour's ghost cells.  The contract is diffChangePar's: bit-identical to the unsharded run and
to the oracle.  On a one-GPU box the slabs share the device (the exchange logic is the same; the stores
just do not cross NVLink); with two GPUs they are spread over both."""
import numpy as np
import pytest

import oracle
import eutropia_b200 as eb
from eutropia_b200 import _lib

pytestmark = pytest.mark.gpu


def _run(m, conc, niter, indvar, devices, rounds=1):
    sf = eb.SlabField(m.mat_in, m.mat_out, conc.shape[1], devices)
    inf = sf.info
    assert inf.n_slabs == len(devices) and inf.n_points == m.nfields
    own = sum(sf.part(i)[2] for i in range(len(devices)))
    assert own == m.nfields                                      # every point has exactly one owner
    sf.upload(conc)
    for _ in range(rounds):
        sf.diffuse(niter, indvar)
    out = sf.download()
    info = sf.info
    sf.close()
    return out, info


@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0, 0, 0]])
def test_slabs_on_one_device_match_oracle(harbor_mesh, devices):
    m = harbor_mesh
    rng = np.random.default_rng(31)
    conc = np.asfortranarray(rng.uniform(0, 25, (m.nfields, 6)) * 10.0 ** rng.integers(-5, 4, (m.nfields, 6)))
    niter, indvar = [14, 0, 5, 14, 9], [6, 2, 1, 3, 4]           # column 5 never listed, column 2 with zero substeps
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 2)
    got, info = _run(m, conc, niter, indvar, devices)
    assert np.array_equal(got, ref)
    if len(devices) > 1:
        assert info.n_links >= 2 * (len(devices) - 1) and info.halo_points > 0 and info.pushes > 0
    # a second round on the resident slabs == two rounds of the oracle
    got2, _ = _run(m, conc, niter, indvar, devices, rounds=2)
    assert np.array_equal(got2, oracle.diffChangePar(m.mat_in, m.mat_out, ref, niter, indvar, 2))


def test_slabs_match_the_unsharded_field_on_a_large_mesh():
    from eutropia_b200 import synth
    m = synth.config_mesh("harbor:300000")
    conc = synth.synthetic_concentrations(m.field_pts, 8, seed=32)
    niter, indvar = np.full(8, 12), np.arange(1, 9)
    plan = eb.Plan(m.mat_in, m.mat_out)
    f = eb.Field(plan, 8)
    f.upload(conc)
    f.diffuse(niter, indvar)
    whole = f.download()
    f.close(); plan.close()
    n_dev = _lib.lib().eut_device_count()
    got, info = _run(m, conc, niter, indvar, [k % n_dev for k in range(4)])
    assert np.array_equal(got, whole)
    assert info.segmented == 1


def test_two_column_groups_pipeline_matches_oracle(harbor_mesh):
    """From 32 columns on, the slab driver runs two column groups as independent substep -> halo ->
    substep pipelines on two streams per slab; mixed substep counts make the groups shrink at different times."""
    m = harbor_mesh
    rng = np.random.default_rng(34)
    c = 40
    conc = np.asfortranarray(rng.uniform(0, 25, (m.nfields, c)))
    niter = rng.integers(0, 9, c).astype(np.int32)
    niter[:4] = [8, 8, 1, 0]
    indvar = rng.permutation(c).astype(np.int32) + 1
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 4)
    n_dev = _lib.lib().eut_device_count()
    got, info = _run(m, conc, niter, indvar, [k % n_dev for k in range(3)])
    assert np.array_equal(got, ref)
    got, _ = _run(m, conc, niter, indvar, [0], rounds=1)
    assert np.array_equal(got, ref)
    # the same call three times: plain launches, then captured as one graph over all slabs and streams, then replayed
    ref3 = ref
    for _ in range(2):
        ref3 = oracle.diffChangePar(m.mat_in, m.mat_out, ref3, niter, indvar, 4)
    got3, info3 = _run(m, conc, niter, indvar, [k % n_dev for k in range(3)], rounds=3)
    assert np.array_equal(got3, ref3)


def test_slabs_on_two_devices_match_oracle(harbor_mesh):
    if _lib.lib().eut_device_count() < 2:
        pytest.skip("needs two GPUs")
    m = harbor_mesh
    rng = np.random.default_rng(33)
    conc = np.asfortranarray(rng.uniform(0, 25, (m.nfields, 5)))
    niter, indvar = [11, 7, 11, 3, 11], [1, 2, 3, 4, 5]
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 2)
    for devices in ([0, 1], [1, 0, 1]):
        got, _ = _run(m, conc, niter, indvar, devices)
        assert np.array_equal(got, ref)


def test_slab_errors_mirror_the_field(harbor_mesh):
    m = harbor_mesh
    sf = eb.SlabField(m.mat_in, m.mat_out, 3, [0, 0])
    with pytest.raises(eb.EutropiaIndexError):
        sf.diffuse([2], [4])                                     # column outside 1..C
    with pytest.raises(eb.EutropiaError):
        sf.diffuse([2, 2], [1, 1])                               # duplicate column
    sf.diffuse([0], [99])                                        # never dereferenced with zero substeps
    sf.close()
    with pytest.raises(eb.EutropiaError):
        eb.SlabField(m.mat_in, m.mat_out, 3, [0, 77])


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0, 0], "all"])
def test_cells_on_slabs_match_the_unsharded_round(devices):
    """eut_slab_cells_*: every slab holds the entries of every cell on its OWN points; per-field work (claim
    rescale, scatter sums in cell order, clamp) is the owner's, per-cell quantities (production shares,
    accessible amounts) are combined over the slabs.  Lists: the union over the slabs is the unsharded list,
    entry for entry, and what only depends on the field point (acc.prop after the claim rescale) has the
    same bits.  A whole round -- gather, scatter + clamp, substeps, gather -- stays within 1e-12 of the
    unsharded field (cells across a cut add their partial sums in another order), two rounds in a row."""
    from eutropia_b200 import synth
    if devices == "all":                      # one slab per GPU of the box (two on a one-GPU box): partial sums cross NVLink
        n_dev = _lib.lib().eut_device_count()
        devices = [k % n_dev for k in range(max(2, n_dev))]
    m = eb.mesh.make_mesh(eb.mesh.harbor_polygon(90.0), 1.0, 3)
    n_cols, n_cells = 6, 220
    conc = synth.synthetic_concentrations(m.field_pts, n_cols, seed=41)
    cx, cy, size, scv = synth.place_cells(m, n_cells, seed=42)
    ex_ptr, ex_cpd, ex_flx = synth.synthetic_exchanges(n_cells, n_cols, 3, seed=43)
    ex_flx = ex_flx * 1e-3
    is_const = np.array([0, 0, 1, 0, 0, 0], dtype=np.uint8)
    niter, indvar = [9, 4, 6, 9, 1], [1, 2, 4, 5, 6]
    # unsharded
    plan = eb.Plan(m.mat_in, m.mat_out)
    fld = eb.Field(plan, n_cols)
    fld.upload(conc)
    cells = eb.Cells(plan)
    cells.build_lists(m.field_pts, cx, cy, size, scv)
    cells.claim_rescale()
    ptr, fid, fd, acc = cells.get_lists()
    # sharded
    sf = eb.SlabField(m.mat_in, m.mat_out, n_cols, devices)
    sf.upload(conc)
    sc = eb.SlabCells(sf)
    sc.build_lists(m.field_pts, cx, cy, size, scv)
    sc.claim_rescale()
    assert sc.n_entries == fid.size
    key = lambda p, f: (np.repeat(np.arange(n_cells), np.diff(p)).astype(np.int64) << 32) | f.astype(np.int64)
    parts = [sc.get_lists(r) for r in range(len(devices))]
    k_all = np.concatenate([key(p[0], p[1]) for p in parts])
    order = np.argsort(k_all, kind="stable")
    ref_order = np.argsort(key(ptr, fid), kind="stable")
    assert np.array_equal(k_all[order], key(ptr, fid)[ref_order])                       # same (cell, field) pairs, each once
    assert np.array_equal(np.concatenate([p[2] for p in parts])[order], fd[ref_order])     # distances: same bits
    assert np.array_equal(np.concatenate([p[3] for p in parts])[order], acc[ref_order])    # acc.prop after the claim rescale: same bits
    assert sum(1 for p in parts if p[1].size) == len(devices)                            # every slab serves some cells
    for _ in range(2):
        a0, b0 = cells.gather(fld, m.field_vol), sc.gather(m.field_vol)
        assert np.all(np.abs(a0 - b0) <= 1e-12 * np.abs(a0))
        cells.scatter(fld, m.field_vol, is_const, ex_ptr, ex_cpd, ex_flx)
        sc.scatter(m.field_vol, is_const, ex_ptr, ex_cpd, ex_flx)
        whole, shard = fld.download(), sf.download()
        assert np.all(np.abs(whole - shard) <= 1e-12 * np.abs(whole)) and np.array_equal(whole == 0, shard == 0)
        fld.diffuse(niter, indvar)
        sf.diffuse(niter, indvar)
        whole, shard = fld.download(), sf.download()
        assert np.all(np.abs(whole - shard) <= 1e-12 * np.abs(whole))
    # a scatter without a current gather takes one itself (the colSums must span the slabs)
    cells.scatter(fld, m.field_vol, is_const, ex_ptr, ex_cpd, ex_flx)
    sc.scatter(m.field_vol, is_const, ex_ptr, ex_cpd, ex_flx)
    whole, shard = fld.download(), sf.download()
    assert np.all(np.abs(whole - shard) <= 1e-12 * np.abs(whole))
    sc.close(); sf.close()


@pytest.mark.parametrize("depth", [2, 3, 5])
def test_deep_halo_exchanges_every_k_substeps(harbor_mesh, depth, monkeypatch):
    """EUT_SLAB_DEPTH = K: every slab keeps K rings of ghost points, the inner K - 1 of them live (diffused like own
    points), and the slabs exchange -- and wait for each other -- only every K-th substep.  The owned points still
    sum the same values in the same order: bit-identical to the oracle, mixed substep counts (not multiples of
    K), two rounds on the resident slabs, three and five slabs on one device."""
    monkeypatch.setenv("EUT_SLAB_DEPTH", str(depth))
    m = harbor_mesh
    rng = np.random.default_rng(33)
    conc = np.asfortranarray(rng.uniform(0, 25, (m.nfields, 6)) * 10.0 ** rng.integers(-5, 4, (m.nfields, 6)))
    niter, indvar = [14, 0, 5, 13, 9], [6, 2, 1, 3, 4]
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 2)
    ref2 = oracle.diffChangePar(m.mat_in, m.mat_out, ref, niter, indvar, 2)
    for devices in ([0, 0, 0], [0, 0, 0, 0, 0]):
        got, info = _run(m, conc, niter, indvar, devices)
        assert np.array_equal(got, ref)
        got2, info2 = _run(m, conc, niter, indvar, devices, rounds=2)
        assert np.array_equal(got2, ref2)
        assert info.halo_points > 0
