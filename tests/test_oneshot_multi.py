"""One call on several GPUs (eut_diffchange_par_multi) and the host side of the one-shot tier:
how the columns are dealt out (CPU), pageable vs pinned caller memory, exact graph recognition."""
import numpy as np
import pytest

import oracle
import eutropia_b200 as eb
from eutropia_b200 import _lib


def deal(niter, indvar, n_devices):
    ni, iv = np.asarray(niter, np.int32), np.asarray(indvar, np.int32)
    own = np.zeros(max(ni.size, 1), np.int32)
    _lib.check(_lib.lib().eut_deal_columns(_lib.ptr(ni), _lib.ptr(iv), ni.size, n_devices, _lib.ptr(own)))
    return own[:ni.size]


def test_columns_are_dealt_in_contiguous_balanced_blocks():
    assert deal([60] * 8, range(1, 9), 4).tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    assert deal([60] * 8, range(1, 9), 1).tolist() == [0] * 8
    # unsorted indvar, a column that does not diffuse: blocks follow the column number
    assert deal([60, 22, 22, 0, 60, 10, 1], [5, 1, 2, 3, 7, 9, 4], 3).tolist() == [1, 0, 0, -1, 2, 2, 0]
    # more devices than columns: one column each, the rest idle
    assert deal([5, 5], [2, 1], 8).tolist() == [1, 0]
    # one heavy column does not starve the others of a device
    assert deal([1000, 1, 1, 1, 1, 1], range(1, 7), 2).tolist() == [0, 1, 1, 1, 1, 1]
    rng = np.random.default_rng(0)
    for _ in range(50):
        n, nd = int(rng.integers(1, 40)), int(rng.integers(1, 9))
        ni = rng.integers(0, 70, n)
        iv = rng.permutation(64)[:n] + 1
        own = deal(ni, iv, nd)
        act = ni > 0
        assert np.all(own[~act] == -1)
        if not act.any():
            continue
        o = own[act][np.argsort(iv[act])]
        assert o[0] == 0 and np.all(np.diff(o) >= 0) and np.all(np.diff(o) <= 1)      # contiguous, none skipped
        assert o[-1] == min(nd, int(act.sum())) - 1                                    # every usable device has work
        work = np.bincount(own[act], weights=ni[act])
        assert work.max() <= ni[act].sum() / len(work) + ni.max()                      # within one column of the mean


def _case(m, rng, n_cols=7):
    conc = rng.uniform(0, 25, (m.nfields, n_cols)) * 10.0 ** rng.integers(-6, 4, (m.nfields, n_cols))
    conc[rng.uniform(size=conc.shape) < 0.05] = 0.0
    niter = [12, 0, 5, 12, 3, 9]
    indvar = [6, 2, 1, 7, 3, 4]              # unsorted; column 5 never listed, column 2 with zero substeps
    return np.asfortranarray(conc), niter, indvar


@pytest.mark.gpu
def test_multi_entry_on_one_device_matches_oracle(harbor_mesh):
    m = harbor_mesh
    conc, niter, indvar = _case(m, np.random.default_rng(21))
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 2)
    got = eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, devices=[0])
    assert np.array_equal(got, ref)
    got = eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, devices="all")    # every visible device
    assert np.array_equal(got, ref)
    with pytest.raises(eb.EutropiaError):
        eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, devices=[0, 0])
    with pytest.raises(eb.EutropiaError):
        eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, devices=[99])


@pytest.mark.gpu
def test_multi_entry_on_two_devices_matches_oracle(harbor_mesh):
    if _lib.lib().eut_device_count() < 2:
        pytest.skip("needs two GPUs")
    m = harbor_mesh
    rng = np.random.default_rng(22)
    conc, niter, indvar = _case(m, rng)
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 2)
    for devs in ([0, 1], [1, 0], [1]):
        for _ in range(3):                   # first call builds the plans, the next ones speculate on them
            got = eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, devices=devs)
            assert np.array_equal(got, ref)
    wide = np.asfortranarray(rng.uniform(0, 25, (m.nfields, 40)))
    ni, iv = np.full(40, 7, np.int32), np.arange(1, 41, dtype=np.int32)
    assert np.array_equal(eb.diffChangePar(m.mat_in, m.mat_out, wide, ni, iv, devices="all"),
                          oracle.diffChangePar(m.mat_in, m.mat_out, wide, ni, iv, 4))


@pytest.mark.gpu
@pytest.mark.parametrize("slot_mb", ["1", "8"])
def test_pageable_and_pinned_caller_memory_agree(harbor_mesh, slot_mb, monkeypatch):
    """numpy arrays are pageable (staged through pinned slot rings by copier threads, several slots per
    chunk at 1 MB); torch pinned tensors take the direct DMA path; EUT_NO_STAGING hands pageable memory to
    the driver as the first version did.  All three must give the oracle's bits."""
    torch = pytest.importorskip("torch")
    monkeypatch.setenv("EUT_STAGE_SLOT_MB", slot_mb)
    monkeypatch.setenv("EUT_COPY_THREADS", "3")
    _lib.lib().eut_oneshot_reset()
    m = harbor_mesh
    rng = np.random.default_rng(23)
    n, c = m.nfields, 24
    conc = np.asfortranarray(rng.uniform(0, 25, (n, c)))
    niter, indvar = np.full(c, 6, np.int32), np.arange(1, c + 1, dtype=np.int32)
    niter[::5] = 11
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 4)
    for _ in range(2):
        assert np.array_equal(eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar), ref)
    pin_in = torch.from_numpy(np.ascontiguousarray(conc.T)).pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    adj = np.require(m.mat_in, dtype=np.int32, requirements=["F", "A"])
    nn = np.require(m.mat_out, dtype=np.int32, requirements=["F", "A"])
    L = _lib.lib()
    for _ in range(2):
        pin_out.zero_()
        _lib.check(L.eut_diffchange_par(_lib.ptr(adj), adj.shape[0], _lib.ptr(nn), nn.shape[0], pin_in.data_ptr(), n, c,
                                        _lib.ptr(niter), _lib.ptr(indvar), c, 1, pin_out.data_ptr(), 0))
        assert np.array_equal(pin_out.numpy().T, ref)
    monkeypatch.setenv("EUT_NO_STAGING", "1")
    assert np.array_equal(eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar), ref)
    monkeypatch.delenv("EUT_NO_STAGING")
    monkeypatch.setenv("EUT_FORCE_STAGED", "1")           # pinned memory through the staged path as well
    pin_out.zero_()
    _lib.check(L.eut_diffchange_par(_lib.ptr(adj), adj.shape[0], _lib.ptr(nn), nn.shape[0], pin_in.data_ptr(), n, c,
                                    _lib.ptr(niter), _lib.ptr(indvar), c, 1, pin_out.data_ptr(), 0))
    assert np.array_equal(pin_out.numpy().T, ref)
    _lib.lib().eut_oneshot_reset()


@pytest.mark.gpu
def test_graph_is_recognised_by_content_not_by_address(small_mesh):
    """Same arrays at new addresses reuse the plan; a graph that differs in ONE entry (same sizes, same
    address as the call before) must not: the byte comparison decides, not a hash."""
    m = small_mesh
    rng = np.random.default_rng(24)
    conc = np.asfortranarray(rng.uniform(0, 25, (m.nfields, 4)))
    niter, indvar = [9, 9, 4, 4], [1, 2, 3, 4]
    mat_in = np.array(m.mat_in, dtype=np.int32, order="F")
    ref = oracle.diffChangePar(mat_in, m.mat_out, conc, niter, indvar, 2)
    for _ in range(2):
        assert np.array_equal(eb.diffChangePar(mat_in.copy(order="F"), m.mat_out, conc, niter, indvar), ref)
    keep = eb.rcpp_api._matrix  # noqa: F841  (the binding passes int32 F-ordered arrays through without a copy)
    adj = np.require(mat_in, dtype=np.int32, requirements=["F", "A"])
    assert np.array_equal(eb.diffChangePar(adj, m.mat_out, conc, niter, indvar), ref)
    assert np.array_equal(eb.diffChangePar(adj, m.mat_out, conc, niter, indvar), ref)      # speculation on the same address
    e = adj.shape[0] // 2
    adj[e, 0] = adj[e, 0] % m.nfields + 1                                                  # rewire one edge in place
    ref2 = oracle.diffChangePar(adj, m.mat_out, conc, niter, indvar, 2)
    assert not np.array_equal(ref2, ref)
    assert np.array_equal(eb.diffChangePar(adj, m.mat_out, conc, niter, indvar), ref2)
