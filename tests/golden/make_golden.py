"""Generates tests/golden/*.npz.  The reference has no golden vectors and cannot be built here
(no R / Rcpp / Armadillo), so these fixtures come from the CPU oracle (oracle/) -- "parity unpinned".
They pin the oracle against regressions and give the GPU tests inputs that do not depend on the
mesh builder.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from eutropia_b200 import mesh, synth  # noqa: E402

here = os.path.dirname(os.path.abspath(__file__))

# 1. diffusion on a small non-convex 4-layer mesh, mixed substep counts, a constant-like column
m = mesh.make_mesh(mesh.harbor_polygon(26.0), 1.0, 4)
conc = synth.synthetic_concentrations(m.field_pts, 6, seed=11)
conc[:, 4] = 3.5
niter = np.array([22, 60, 7, 1], dtype=np.int32)
indvar = np.array([1, 3, 6, 2], dtype=np.int32)
expected = oracle.diffChangePar(m.mat_in, m.mat_out, conc, niter, indvar, 2)
np.savez_compressed(os.path.join(here, "diffusion_small.npz"), mat_in=m.mat_in, mat_out=m.mat_out,
                    conc=conc, niter=niter, indvar=indvar, expected=expected, field_pts=m.field_pts)

# 2. cell exchange on the same mesh
cx, cy, size, scv = synth.place_cells(m, 40, seed=12)
ptr, fid, fd, acc = oracle.cellmap(m.field_pts, cx, cy, size, scv)
acc2 = oracle.claim(fid, acc, m.nfields)
fmol = oracle.gather(conc, ptr, fid, acc2, m.field_vol)
ex_ptr, ex_cpd, ex_flx = synth.synthetic_exchanges(40, 6, 3, seed=13)
is_const = np.array([0, 0, 0, 0, 1, 0], dtype=np.uint8)
after = oracle.scatter(conc, is_const, ptr, fid, fd, acc2, m.field_vol, ex_ptr, ex_cpd, ex_flx)
np.savez_compressed(os.path.join(here, "cells_small.npz"), field_pts=m.field_pts, mat_in=m.mat_in,
                    mat_out=m.mat_out, conc=conc, cx=cx, cy=cy, size=size, scv=scv, cell_ptr=ptr,
                    field_id=fid, field_dist=fd, acc_raw=acc, acc_prop=acc2, fmol=fmol,
                    ex_ptr=ex_ptr, ex_cpd=ex_cpd, ex_flx=ex_flx, is_const=is_const, after=after,
                    field_vol=m.field_vol)
print("golden fixtures written:", m.nfields, "points,", fid.size, "cell entries")
