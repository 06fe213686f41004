"""The pieces of a round either side of the stencil (SURVEY 8a-4', 8a-7, 8f-1): exoenzyme catalysis
and decay, the legacy `diffuse_compounds` dispatcher, `adjust_uptake`.  CPU part: the oracle's
Lambert W, R's `cut` chunking, the LP-bound arithmetic.  GPU part: device kernels against the oracle
(1e-12 relative, the bar of BASELINE.json for floating point)."""
import math

import numpy as np
import pytest

import oracle
from eutropia_b200 import environment, exchange, mesh

TOL = 1e-12


def test_oracle_lambert_w0_known_values():
    # W(0) = 0, W(e) = 1, W(1) = Omega, W(-1/e) = -1; and the defining identity W e^W = x
    assert oracle.lambertW0(0.0) == 0.0
    assert abs(oracle.lambertW0(math.e) - 1.0) < 1e-15
    assert abs(oracle.lambertW0(1.0) - 0.5671432904097838) < 1e-15
    assert abs(oracle.lambertW0(-1.0 / math.e) + 1.0) < 1e-7
    x = np.concatenate([np.logspace(-300, 300, 301), -np.logspace(-12, np.log10(0.36), 40)])
    w = oracle.lambertW0(x)
    assert np.all(np.abs(w * np.exp(w) - x) <= 4e-16 * np.abs(x) * (1 + np.abs(w)))
    assert np.isinf(oracle.lambertW0(np.inf)) and np.isnan(oracle.lambertW0(-1.0))


def test_chunk_codes_match_r_cut():
    # cut(1:10, 3, labels = FALSE) -> 1 1 1 1 2 2 2 3 3 3 ; cut(1:7, 2) -> 1 1 1 1 2 2 2 ; cut(1:5, 5) -> 1..5
    assert environment.chunk_codes(10, 3).tolist() == [1, 1, 1, 1, 2, 2, 2, 3, 3, 3]
    assert environment.chunk_codes(7, 2).tolist() == [1, 1, 1, 1, 2, 2, 2]
    assert environment.chunk_codes(5, 5).tolist() == [1, 2, 3, 4, 5]
    assert environment.chunk_codes(1, 1).tolist() == [1]


def test_adjust_uptake_arithmetic():
    react_id = ["BIO", "EX_glc", "R1", "EX_o2", "EX_ac"]
    lowbnd = np.array([0.0, -10.0, -1000.0, -1000.0, 0.0])
    acc = {"compounds": ["o2", "xyz", "glc", "ac"], "fmol": np.array([3.0, 9.0, 0.5, 7.0])}
    ind, lb = exchange.adjust_uptake(react_id, lowbnd, cMass=0.28, accCpdFMOL=acc, deltaTime=1 / 6)
    assert ind.tolist() == [4, 2, 5]                       # order of the compounds, 1-based reaction ids
    want = -np.minimum(np.array([3.0, 0.5, 7.0]) * (1 / (1 / 6)) / 0.28, -lowbnd[[3, 1, 4]])
    assert np.array_equal(lb, want)
    assert lb[2] == 0.0                                    # lowbnd 0: no uptake whatever is around
    assert exchange.adjust_uptake(["BIO"], [0.0], 0.28, acc, 1 / 6) == (None, None)


def _small_env(layers=3, seed=3, polygon="Rectangle_30_24", **kw):
    m = mesh.make_mesh(polygon, 1.0, layers)
    rng = np.random.default_rng(seed)
    cpds = ["glc", "ac", "o2", "sucr", "fru", "const"]
    conc = rng.uniform(0, 25, (m.nfields, len(cpds)))
    conc[rng.uniform(size=conc.shape) < 0.1] = 0.0
    conc[:, 5] = 1000.0
    return m, cpds, conc, rng


@pytest.mark.gpu
def test_exoenzyme_catalysis_and_decay_match_oracle():
    m, cpds, conc, rng = _small_env()
    exo = rng.uniform(0, 50, (m.nfields, 2))
    exo[rng.uniform(size=exo.shape) < 0.3] = 0.0
    enzymes = [environment.Exoenzyme("inv", ["sucr", "glc", "fru"], [-1.0, 1.0, 1.0], Kcat=1200.0, Km=4.0, D=10.0, lambda_=0.35),
               environment.Exoenzyme("x", ["glc", "const", "ac"], [-1.0, 2.0, 0.5], Kcat=3e5, Km=0.8, D=0.0, lambda_=0.0)]
    is_const = np.array([False] * 5 + [True])
    env = environment.GrowthEnvironment(m, cpds, conc, np.full(6, 75.0), is_const, exoenzymes=enzymes, exoenzymes_conc=exo)
    dt = 1 / 6
    env.exoenzyme_activity(dt)
    want_c, want_e = conc.copy(), exo.copy()
    for k, ex in enumerate(enzymes, start=1):
        mets = [cpds.index(x) + 1 for x in ex.mets]
        want_c = oracle.exoenzyme_catalysis(want_c, want_e, k, ex.Kcat, ex.Km, mets, ex.stoich, is_const, dt)
    for k, ex in enumerate(enzymes, start=1):
        want_e = oracle.exoenzyme_decay(want_e, k, ex.lambda_, dt)
    got_c, got_e = env.concentrations, env.exoenzymes_conc
    assert np.array_equal(got_c[:, 5], conc[:, 5])                       # constant compound untouched
    # dS = S_0 - S_t cancels when little is consumed: the bar is 1e-12 of the substrate scale of the
    # point (a product column that starts at 0 receives dS itself, whose last bits depend on libm)
    scale = np.maximum(np.abs(want_c), np.abs(conc[:, :5]).max(axis=1, keepdims=True))
    err = np.abs(got_c - want_c) / scale
    assert np.all(np.isfinite(got_c)) and err.max() <= TOL, err.max()
    assert np.array_equal(got_e, want_e)                                 # one multiplication: bit-exact
    assert np.any(want_c[:, 3] < conc[:, 3]) and np.all(want_c[:, 3] >= 0)   # substrate is consumed, never negative


@pytest.mark.gpu
def test_exoenzyme_decay_waits_for_catalysis_on_a_large_mesh():
    """Catalysis runs on the concentration field's stream and READS the exoenzyme column; the decay
    runs on the exoenzyme field's stream and overwrites it (R/run_simulation.R:343-371).  On a mesh
    large enough that the catalysis kernel is still running when the decay is queued, the second
    must wait for the first (an event between the two streams) -- several rounds back to back."""
    m, cpds, conc, rng = _small_env(polygon="Rectangle_520_500", seed=11)
    exo = rng.uniform(1, 50, (m.nfields, 1))
    enz = [environment.Exoenzyme("inv", ["sucr", "glc", "fru"], [-1.0, 1.0, 1.0], Kcat=1200.0, Km=4.0, D=10.0, lambda_=0.9)]
    env = environment.GrowthEnvironment(m, cpds, conc, np.full(6, 75.0), None, exoenzymes=enz, exoenzymes_conc=exo)
    dt = 1 / 6
    want_c, want_e = conc.copy(), exo.copy()
    for _ in range(3):
        env.exoenzyme_activity(dt)
        want_c = oracle.exoenzyme_catalysis(want_c, want_e, 1, 1200.0, 4.0, [4, 1, 5], enz[0].stoich, None, dt)
        want_e = oracle.exoenzyme_decay(want_e, 1, 0.9, dt)
    got_c, got_e = env.concentrations, env.exoenzymes_conc
    assert np.array_equal(got_e, want_e)
    scale = np.maximum(np.maximum(np.abs(want_c), np.abs(conc[:, :5]).max(axis=1, keepdims=True)), 1e-300)   # rows of zeros stay zeros
    assert np.all(np.isfinite(got_c)) and (np.abs(got_c - want_c) / scale).max() <= 3 * TOL


@pytest.mark.gpu
def test_legacy_dispatcher_equals_par_dispatcher_and_oracle():
    m, cpds, conc, _ = _small_env(seed=5)
    D = np.array([75.0, 10.0, 0.0, 200.0, 75.0, 75.0])
    dt = 1 / 6
    a = environment.GrowthEnvironment(m, cpds, conc, D)
    b = environment.GrowthEnvironment(m, cpds, conc, D)
    a.diffuse_compoundsPar(dt, n_cores=4)
    b.diffuse_compounds(dt, n_cores=4)                    # chunks of columns through serial diffChange
    ind = np.array([1, 2, 4, 5])                          # sd > 0 & D > 0 (column 6 is uniform, column 3 has D = 0)
    niter = environment.diffusion_niter(D[ind - 1], dt, m.field_size)
    want = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter.astype(np.int32), ind.astype(np.int32), 2)
    assert np.array_equal(a.concentrations, want)
    assert np.array_equal(b.concentrations, want)
