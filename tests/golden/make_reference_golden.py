"""Generates tests/golden/diffusion_reference.npz from the REFERENCE ITSELF: src/diffChange.cpp compiled
unmodified into oracle/_ref (oracle/Makefile target `ref`, Armadillo stand-in oracle/ref_shim/).  Needs
/root/reference, so it runs in the build container only; the vectors travel to the GPU box.
Run from the repo root:  python tests/golden/make_reference_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from eutropia_b200 import mesh, synth  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))
assert oracle.ref_build(force=True), "oracle/_ref did not build"
out = {}

# case A: diffChangePar on a non-convex 5-layer mesh, unordered indvar, zero and large substep counts
m = mesh.make_mesh(mesh.harbor_polygon(30.0), 1.0, 5)
conc = synth.synthetic_concentrations(m.field_pts, 7, seed=21)
conc[:, 5] = 0.25
niter = np.array([13, 0, 41, 2, 97], dtype=np.int32)
indvar = np.array([7, 2, 1, 6, 4], dtype=np.int32)
out.update(a_mat_in=m.mat_in, a_mat_out=m.mat_out, a_conc=conc, a_niter=niter, a_indvar=indvar,
           a_expected=oracle.ref_diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 3))

# case B: serial diffChange (every argument double), niter shorter than the column count, and the
# `ychg = ychg * 0` carry-over: an Inf in column 1 and a NaN in column 2 leak into later columns
m = mesh.make_mesh("Petri_17", 1.0, 4)
conc = synth.synthetic_concentrations(m.field_pts, 5, seed=22)
conc[5, 0] = np.inf
conc[40, 1] = np.nan
niter = np.array([2.0, 3.0, 2.0, 1.0])
out.update(b_mat_in=m.mat_in.astype(np.float64), b_mat_out=m.mat_out.astype(np.float64), b_conc=conc, b_niter=niter,
           b_expected=oracle.ref_diffChange(m.mat_in, m.mat_out, conc, niter))

# case C: finite serial case on a single-layer rectangle with fractional substep counts (k < niter(i))
m = mesh.make_mesh("Rectangle_25_12", 1.0, 1)
conc = synth.synthetic_concentrations(m.field_pts, 3, seed=23)
niter = np.array([2.5, 0.2, 6.0])
out.update(c_mat_in=m.mat_in, c_mat_out=m.mat_out, c_conc=conc, c_niter=niter,
           c_expected=oracle.ref_diffChange(m.mat_in, m.mat_out, conc, niter))

# the older oracle-made fixture must agree with the reference as well
g = np.load(os.path.join(here, "diffusion_small.npz"))
assert np.array_equal(oracle.ref_diffChangePar(g["mat_in"], g["mat_out"], g["conc"], g["niter"], g["indvar"], 2),
                      g["expected"]), "diffusion_small.npz disagrees with the compiled reference"

np.savez_compressed(os.path.join(here, "diffusion_reference.npz"), **out)
print("reference golden vectors written:", {k: v.shape for k, v in out.items() if k.endswith("expected")})
