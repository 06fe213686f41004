"""GPU parity tests of the diffusion path, through the C ABI (one-shot and session tiers), against
the CPU oracle, against vectors the reference itself produced (tests/golden/diffusion_reference.npz)
and against the reference's compiled src/diffChange.cpp (oracle/_ref) where that library is present.  Bar: bit-exact (np.array_equal) on finite inputs -- stricter than the 1e-12
relative tolerance BASELINE.json asks for, which is also asserted explicitly (REL_TOL)."""
import os

import numpy as np
import pytest

import oracle
import eutropia_b200 as eb
from eutropia_b200 import _lib, mesh, synth

pytestmark = pytest.mark.gpu
REL_TOL = 1e-12
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def assert_parity(got, ref):
    assert got.shape == ref.shape
    same = (got == ref) | (np.isnan(got) & np.isnan(ref))       # +-Inf compare equal to themselves
    with np.errstate(invalid="ignore"):
        err = np.abs(got - ref)
        bound = REL_TOL * np.maximum(np.abs(ref), 1e-300)
        assert np.all((err <= bound) | same), float(np.nanmax(err / bound))
    assert np.all(same), f"{np.count_nonzero(~same)} entries differ in the last bits"


def test_golden_fixture_one_shot():
    g = np.load(os.path.join(GOLD, "diffusion_small.npz"))
    out = eb.diffChangePar(g["mat_in"], g["mat_out"], g["conc"], g["niter"], g["indvar"], 4)
    assert_parity(out, g["expected"])


def test_reference_golden_vectors_one_shot():
    """Vectors written by the reference's own code (tests/golden/make_reference_golden.py)."""
    g = np.load(os.path.join(GOLD, "diffusion_reference.npz"))
    assert_parity(eb.diffChangePar(g["a_mat_in"], g["a_mat_out"], g["a_conc"], g["a_niter"], g["a_indvar"], 8),
                  g["a_expected"])
    # serial entry point, non-finite data: the `ychg = ychg * 0` carry-over between columns
    assert_parity(eb.diffChange(g["b_mat_in"], g["b_mat_out"], g["b_conc"], g["b_niter"]), g["b_expected"])
    assert_parity(eb.diffChange(g["c_mat_in"], g["c_mat_out"], g["c_conc"], g["c_niter"]), g["c_expected"])


def test_reference_golden_vectors_session():
    g = np.load(os.path.join(GOLD, "diffusion_reference.npz"))
    plan = eb.Plan(g["a_mat_in"], g["a_mat_out"])
    fld = eb.Field(plan, g["a_conc"].shape[1])
    fld.upload(g["a_conc"])
    fld.diffuse(g["a_niter"], g["a_indvar"])
    assert_parity(fld.download(), g["a_expected"])


@pytest.mark.skipif(not oracle.ref_available(), reason="oracle/_ref/libeutropia_ref.so did not travel")
def test_cuda_matches_compiled_reference():
    """The CUDA path against src/diffChange.cpp itself on a mesh large enough for full tiles."""
    m = mesh.make_mesh(mesh.harbor_polygon(150.0), 1.0, 4)
    conc = synth.synthetic_concentrations(m.field_pts, 8, seed=5)
    niter = np.array([9, 30, 1, 0, 14, 3], dtype=np.int32)
    indvar = np.array([8, 2, 5, 3, 1, 7], dtype=np.int32)
    ref = oracle.ref_diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, oracle.ref_max_threads())
    assert_parity(eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 1), ref)
    ref = oracle.ref_diffChange(m.mat_in, m.mat_out, conc, [4.0, 0.5, 0.0, 7.0])
    assert_parity(eb.diffChange(m.mat_in, m.mat_out, conc, [4.0, 0.5, 0.0, 7.0]), ref)
    bad = conc[:, :5].copy()
    bad[17, 0] = np.inf; bad[1234, 1] = np.nan; bad[:, 4] = -bad[:, 4]
    ref = oracle.ref_diffChange(m.mat_in, m.mat_out, bad, [3, 2, 0, 4, 1])
    assert np.isnan(ref[:, 3]).any()                       # the carry-over reached a finite column
    assert_parity(eb.diffChange(m.mat_in, m.mat_out, bad, [3, 2, 0, 4, 1]), ref)


@pytest.mark.parametrize("order", [_lib.ORDER_IDENTITY, _lib.ORDER_ROWBLOCK, _lib.ORDER_TILED])
@pytest.mark.parametrize("generic", [False, True])
def test_session_matches_oracle_all_variants(harbor_mesh, order, generic):
    m = harbor_mesh
    conc = synth.synthetic_concentrations(m.field_pts, 9, seed=21)
    niter = [22, 60, 5, 0, 13, 60, 1, 2, 33]
    indvar = [1, 2, 3, 4, 5, 6, 7, 8, 9]
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 4)
    plan = eb.Plan(m.mat_in, m.mat_out, order=order, force_generic=generic)
    assert (plan.info.ell_width == 0) == generic
    assert bool(plan.info.tiled) == (order == _lib.ORDER_TILED and not generic)
    f = eb.Field(plan, 9)
    f.upload(conc)
    f.diffuse(niter, indvar)
    assert_parity(f.download(), ref)
    # a second round on the resident field == two rounds of the oracle
    f.diffuse(niter, indvar)
    ref2 = oracle.diffChangePar(m.mat_in, m.mat_out, ref, niter, indvar, 4)
    assert_parity(f.download(), ref2)


def test_index_map_bit_exact(harbor_mesh):
    m = harbor_mesh
    n = m.nfields
    to, frm = m.mat_in[:, 1].astype(np.int64), m.mat_in[:, 0]
    order = np.argsort(to, kind="stable")
    row_ptr = np.concatenate([[0], np.cumsum(np.bincount(to, minlength=n + 1)[1:])])
    # ORDER_SEGMENTS: the map is replayed from the segment kernel's own tables (meta, copy lists,
    # segment and generic-point records) read back from the device, not from the ELL tables
    for o in (_lib.ORDER_IDENTITY, _lib.ORDER_ROWBLOCK, _lib.ORDER_TILED, _lib.ORDER_SEGMENTS):
        for generic in (False, True):
            plan = eb.Plan(m.mat_in, m.mat_out, order=o, force_generic=generic)
            assert bool(plan.info.segmented) == (o == _lib.ORDER_SEGMENTS and not generic)
            rp, col = plan.csr()
            assert np.array_equal(rp, row_ptr)
            assert np.array_equal(col, frm[order])
            perm = plan.perm()
            assert np.unique(perm).size == n and perm.min() >= 0 and perm.max() < plan.info.n_padded   # injective (segment plans leave gap cells)


def test_shuffled_mat_in_keeps_row_order_semantics(small_mesh):
    # mat.in rows in arbitrary order: the per-destination summation order is the row order
    m = small_mesh
    rng = np.random.default_rng(8)
    sh = rng.permutation(m.mat_in.shape[0])
    mat_in = np.ascontiguousarray(m.mat_in[sh])
    mat_out = np.ascontiguousarray(m.mat_out[rng.permutation(m.nfields)])
    conc = rng.uniform(0, 25, (m.nfields, 3)) * 10.0 ** rng.integers(-8, 8, (m.nfields, 3))
    ref = oracle.diffChangePar(mat_in, mat_out, conc, [9, 4, 1], [3, 1, 2], 2)
    assert_parity(eb.diffChangePar(mat_in, mat_out, conc, [9, 4, 1], [3, 1, 2], 2), ref)


def test_diffchange_double_interface(small_mesh):
    m = small_mesh
    rng = np.random.default_rng(9)
    conc = rng.uniform(0, 25, (m.nfields, 4))
    adj = m.mat_in.astype(np.float64) + 0.25           # (int) truncation, src/diffChange.cpp:20
    nn = m.mat_out.astype(np.float64)
    niter = [3.0, 0.0, 2.5]                            # fewer entries than columns; k < 2.5 -> 3
    ref = oracle.diffChange(adj, nn, conc, niter)
    assert_parity(eb.diffChange(adj, nn, conc, niter), ref)
    assert np.array_equal(ref[:, 3], conc[:, 3])


def test_asymmetric_irregular_graph():
    # random directed graph: in-degree up to 20 (generic path), duplicate / missing mat.out rows
    rng = np.random.default_rng(10)
    n, e = 500, 4000
    mat_in = np.stack([rng.integers(1, n + 1, e), rng.integers(1, n + 1, e)], axis=1).astype(np.int32)
    ids = np.concatenate([rng.permutation(n)[: n - 20] + 1, rng.integers(1, n + 1, 15)]).astype(np.int32)
    mat_out = np.stack([ids, rng.integers(1, 14, ids.size).astype(np.int32)], axis=1)
    conc = rng.uniform(0, 25, (n, 3))
    ref = oracle.diffChangePar(mat_in, mat_out, conc, [6, 3, 1], [1, 2, 3], 2)
    plan = eb.Plan(mat_in, mat_out, n_points=n)
    assert plan.info.ell_width == 0 and plan.info.regular_outflow == 0
    f = eb.Field(plan, 3)
    f.upload(conc)
    f.diffuse([6, 3, 1], [1, 2, 3])
    assert_parity(f.download(), ref)
    assert_parity(eb.diffChangePar(mat_in, mat_out, conc, [6, 3, 1], [1, 2, 3]), ref)


def test_edge_cases(small_mesh):
    m = small_mesh
    conc = np.random.default_rng(11).uniform(0, 25, (m.nfields, 2))
    # nothing to do: empty indvar, zero substeps
    assert np.array_equal(eb.diffChangePar(m.mat_in, m.mat_out, conc, [], []), conc)
    assert np.array_equal(eb.diffChangePar(m.mat_in, m.mat_out, conc, [0, 0], [1, 2]), conc)
    # a bad column with zero substeps is never dereferenced by the reference either
    assert np.array_equal(eb.diffChangePar(m.mat_in, m.mat_out, conc, [0], [7]), conc)
    # no edges at all: only the outflow term acts
    empty = np.zeros((0, 2), dtype=np.int32)
    ref = oracle.diffChangePar(empty, m.mat_out, conc, [2, 1], [1, 2], 1)
    assert_parity(eb.diffChangePar(empty, m.mat_out, conc, [2, 1], [1, 2]), ref)
    # single point, no neighbours
    one_out = np.array([[1, 1]], dtype=np.int32)
    c1 = np.array([[3.0]])
    assert_parity(eb.diffChangePar(empty, one_out, c1, [5], [1]), oracle.diffChangePar(empty, one_out, c1, [5], [1], 1))
    # signed zeros, subnormals, huge values
    sp = conc.copy()
    sp[::7, 0] = -0.0; sp[1::7, 0] = 5e-324; sp[2::7, 0] = 1e300; sp[:, 1] = -0.0
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, sp, [4, 3], [1, 2], 1)
    got = eb.diffChangePar(m.mat_in, m.mat_out, sp, [4, 3], [1, 2])
    assert_parity(got, ref)
    assert np.array_equal(np.signbit(got), np.signbit(ref))
    # serial routine with -0.0 columns (ychg carries its sign across columns in the reference)
    ref = oracle.diffChange(m.mat_in, m.mat_out, sp, [4, 3])
    got = eb.diffChange(m.mat_in, m.mat_out, sp, [4, 3])
    assert_parity(got, ref)
    assert np.array_equal(np.signbit(got), np.signbit(ref))


def test_errors_mirror_reference(small_mesh):
    m = small_mesh
    conc = np.ones((m.nfields, 2))
    bad = m.mat_in.copy(); bad[5, 0] = m.nfields + 1
    with pytest.raises(eb.EutropiaIndexError):
        eb.diffChangePar(bad, m.mat_out, conc, [1], [1])
    bad = m.mat_in.copy(); bad[5, 1] = 0
    with pytest.raises(eb.EutropiaIndexError):
        eb.diffChangePar(bad, m.mat_out, conc, [1], [1])
    with pytest.raises(eb.EutropiaIndexError):
        eb.diffChangePar(m.mat_in, m.mat_out, conc, [1], [3])
    with pytest.raises(eb.EutropiaIndexError):
        eb.diffChangePar(m.mat_in, m.mat_out, conc, [1], [1, 2])      # niter shorter than indvar
    with pytest.raises(eb.EutropiaIndexError):
        eb.diffChange(m.mat_in, m.mat_out, conc, [1, 1, 1])            # niter longer than columns
    with pytest.raises(eb.EutropiaError):
        eb.diffChangePar(m.mat_in, m.mat_out, conc, [1, 1], [1, 1])    # duplicate column
    # inputs are never written
    c0 = conc.copy()
    eb.diffChangePar(m.mat_in, m.mat_out, conc, [3, 3], [1, 2])
    assert np.array_equal(conc, c0)


def test_properties_at_scale():
    # 1M-point class mesh is too slow for the oracle's full run: check size-independent properties
    m = synth.config_mesh("harbor:300000")
    conc = synth.synthetic_concentrations(m.field_pts, 8, seed=31)
    env_niter = [60] * 8
    out = eb.diffChangePar(m.mat_in, m.mat_out, conc, env_niter, list(range(1, 9)))
    assert np.allclose(out.sum(0), conc.sum(0), rtol=1e-11)            # mass conservation
    assert out.min() >= 0                                              # positivity
    assert np.all(out.max(0) <= conc.max(0)) and np.all(out.min(0) >= conc.min(0))  # max principle
    # linearity in the number of substeps: 60 == 25 + 35 (bit-exact: same operation sequence)
    a = eb.diffChangePar(m.mat_in, m.mat_out, conc, [25] * 8, list(range(1, 9)))
    b = eb.diffChangePar(m.mat_in, m.mat_out, a, [35] * 8, list(range(1, 9)))
    assert np.array_equal(b, out)
    # spot-check two columns against the oracle at full size
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc[:, :2], [60, 60], [1, 2], 8)
    assert_parity(out[:, :2], ref)
    # uniform field is a fixed point up to rounding
    uni = np.full((m.nfields, 1), 12.5)
    assert np.allclose(eb.diffChangePar(m.mat_in, m.mat_out, uni, [60], [1]), uni, rtol=1e-13)


def test_dispatcher_diffuse_compoundsPar(harbor_mesh):
    m = harbor_mesh
    n = m.nfields
    rng = np.random.default_rng(41)
    conc = np.empty((n, 6), order="F")
    conc[:, 0] = rng.uniform(0, 25, n)      # D = 75  -> 60 substeps
    conc[:, 1] = 1000.0                     # uniform -> sd == 0 -> skipped
    conc[:, 2] = rng.uniform(0, 5, n)       # constant flag -> skipped
    conc[:, 3] = rng.uniform(0, 5, n)       # D = 0 -> skipped
    conc[:, 4] = rng.uniform(0, 5, n)       # D = Inf -> column mean
    conc[:, 5] = rng.uniform(0, 5, n)       # D = 10 -> 22 substeps
    D = np.array([75.0, 75.0, 75.0, 0.0, np.inf, 10.0])
    const = np.array([0, 0, 1, 0, 0, 0], dtype=bool)
    assert eb.diffusion_niter([75.0, 10.0], 1 / 6, 1.0).tolist() == [60.0, 22.0]
    env = eb.GrowthEnvironment(m, [f"cpd{i}" for i in range(6)], conc, D, const)
    env.diffuse_compoundsPar(1 / 6)
    got = env.concentrations
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, [60, 22], [1, 6], 2)
    ref[:, 4] = conc[:, 4].mean()
    assert np.array_equal(got[:, [0, 1, 2, 3, 5]], ref[:, [0, 1, 2, 3, 5]])
    assert np.allclose(got[:, 4], ref[:, 4], rtol=1e-12)


def test_one_shot_speculation_is_verified(small_mesh):
    """The one-shot tier starts the first column chunks on the plan of the previous call while the
    graph hash is still being taken (same array addresses as last time).  A graph that changed IN
    PLACE must still be detected: the guess is discarded and the result follows the new graph."""
    m = small_mesh
    adj = np.require(m.mat_in.reshape(-1, 2), dtype=np.int32, requirements=["F", "A", "W", "O"]).copy(order="F")
    nn = np.require(m.mat_out.reshape(-1, 2), dtype=np.int32, requirements=["F", "A"]).copy(order="F")
    conc = synth.synthetic_concentrations(m.field_pts, 9, seed=11)
    niter = [7] * 9
    indvar = list(range(1, 10))
    a = eb.diffChangePar(adj, nn, conc, niter, indvar)          # builds and caches the plan
    b = eb.diffChangePar(adj, nn, conc, niter, indvar)          # same addresses: speculated
    ref = oracle.diffChangePar(adj, nn, conc, niter, indvar, 2)
    assert np.array_equal(a, ref) and np.array_equal(b, ref)
    # rewire a handful of edges in place (same buffers, same sizes): drop them onto other sources
    rng = np.random.default_rng(1)
    rows = rng.choice(adj.shape[0], 40, replace=False)
    adj[rows, 0] = rng.integers(1, m.nfields + 1, rows.size)
    c = eb.diffChangePar(adj, nn, conc, niter, indvar)          # guess is wrong -> must fall back
    ref2 = oracle.diffChangePar(adj, nn, conc, niter, indvar, 2)
    assert np.array_equal(c, ref2) and not np.array_equal(c, ref)
    d = eb.diffChangePar(adj, nn, conc, niter, indvar)          # and the new graph is now the guess
    assert np.array_equal(d, ref2)


@pytest.mark.parametrize("mode", ["0", "1"])
def test_one_shot_pipeline_modes(small_mesh, tmp_path, mode):
    """EUT_ONESHOT_STREAM selects the pipeline of the one-shot tier (read once per process, hence the
    subprocess): 0 the chunk ramp on two compute streams (default), 1 column streaming with a host-paced launch
    loop.  Same bits as the oracle from both: mixed substep counts, columns in any order, 37 columns so that groups join and
    leave at different substeps; three calls in a row (the second captures CUDA graphs, the third replays
    them) with the result fed back in, as R does every round; pinned-like and pageable caller memory."""
    import subprocess
    import sys
    m = small_mesh
    rng = np.random.default_rng(23)
    conc = synth.synthetic_concentrations(m.field_pts, 40, seed=23)
    indvar = (rng.permutation(40)[:37] + 1).tolist()
    niter = rng.integers(1, 31, 37).tolist()
    niter[:4] = [30, 1, 2, 17]
    inp, outp = tmp_path / "in.npz", tmp_path / "out.npy"
    np.savez(inp, mat_in=m.mat_in, mat_out=m.mat_out, conc=conc, niter=niter, indvar=indvar)
    code = ("import numpy as np, eutropia_b200 as eb; g = np.load(r'%s'); c = g['conc']\n"
            "for _ in range(3): c = eb.diffChangePar(g['mat_in'], g['mat_out'], c, g['niter'], g['indvar'])\n"
            "np.save(r'%s', c)" % (inp, outp))
    env = dict(os.environ, PYTHONPATH=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env.pop("EUT_ONESHOT_STREAM", None)
    if mode != "0":
        env["EUT_ONESHOT_STREAM"] = mode                     # the library tests for presence, not for the value
    subprocess.run([sys.executable, "-c", code], check=True, env=env, timeout=300)
    ref = conc
    for _ in range(3):
        ref = oracle.diffChangePar(m.mat_in, m.mat_out, ref, niter, indvar, 2)
    assert_parity(np.load(outp), ref)


@pytest.mark.parametrize("chunks", ["1,2,3,1,4", "2,2,2,2,2,1", "5,6"])
def test_one_shot_two_lane_chunk_pipeline(harbor_mesh, chunks, monkeypatch):
    """The one-shot pipeline alternates its column chunks between two compute streams, each chunk with
    its own range of the launch tables.  Small meshes fit one chunk, so the chunk sizes are forced
    (EUT_ONESHOT_CHUNKS is read per call): odd and even substep counts, columns in any order, repeated
    calls (the second and third replay CUDA graphs), and the single-lane setting must all give the
    oracle's bits."""
    m = harbor_mesh
    conc = synth.synthetic_concentrations(m.field_pts, 13, seed=31)
    niter = [9, 3, 12, 1, 7, 7, 2, 30, 5, 4, 8]
    indvar = [1, 3, 2, 5, 4, 6, 8, 13, 11, 9, 10]
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 2)
    monkeypatch.setenv("EUT_ONESHOT_CHUNKS", chunks)
    for _ in range(3):
        assert_parity(eb.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar), ref)
    # feed the result back in, as R does every round
    ref2 = oracle.diffChangePar(m.mat_in, m.mat_out, ref, niter, indvar, 2)
    assert_parity(eb.diffChangePar(m.mat_in, m.mat_out, ref, niter, indvar), ref2)


@pytest.mark.parametrize("n_cols", [40, 70, 131])
def test_session_lanes_match_oracle(harbor_mesh, n_cols, monkeypatch):
    """A wide diffuse call deals its columns out to 2..4 lanes (streams) that run side by side
    (field.cu).  Small meshes stay below the lane threshold, so the lanes are forced here: mixed
    substep counts (zeros, odd, even), columns in any order, three calls in a row (the second captures,
    the third replays the CUDA graph with its fork / join), then a narrow call on the same field."""
    m = harbor_mesh
    monkeypatch.setenv("EUT_DIFFUSE_LANES_FORCE", "1")
    rng = np.random.default_rng(n_cols)
    conc = synth.synthetic_concentrations(m.field_pts, n_cols, seed=n_cols)
    indvar = rng.permutation(n_cols)[: n_cols - 3] + 1
    niter = rng.integers(0, 9, indvar.size)
    niter[:4] = [8, 1, 0, 5]
    plan = eb.Plan(m.mat_in, m.mat_out)
    assert plan.info.segmented == 1
    fld = eb.Field(plan, n_cols)
    fld.upload(conc)
    ref = conc
    for _ in range(3):
        fld.diffuse(niter, indvar)
        ref = oracle.diffChangePar(m.mat_in, m.mat_out, ref, niter, indvar, 4)
        assert_parity(fld.download(), ref)
    fld.diffuse([3, 2], [2, 1])
    ref = oracle.diffChangePar(m.mat_in, m.mat_out, ref, [3, 2], [2, 1], 2)
    assert_parity(fld.download(), ref)
    # the same through the ELL cross-check kernel (never uses lanes)
    fld2 = eb.Field(plan, n_cols)
    fld2.set_option("variant", 1)
    fld2.upload(conc)
    fld2.diffuse(niter, indvar)
    assert_parity(fld2.download(), oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 4))
