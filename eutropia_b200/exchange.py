"""Host-side mirrors of the two small R functions either side of the per-cell LP solve:
`scavenge_compounds` (R/scavenge_compounds.R:12-31) and `adjust_uptake` (R/adjust_uptake.R:9-33).
The LP itself (sybil / GLPK, R/run_simulation.R:630-688) stays on the host solver and is out of scope.
The heavy part -- gathering the accessible amounts of every compound for every cell -- runs on the
device (`Cells.gather`); these functions only give the result the reference's shape."""
from __future__ import annotations

import numpy as np

__all__ = ["scavenge_compounds", "adjust_uptake", "chemotaxis_shift"]


def scavenge_compounds(cells, field, env_cpds, env_fieldVol, cell_index=None):
    """Accessible amount [fmol] of every compound for every cell: `apply(conc[field.id, ] / 1000 *
    fieldVol * acc.prop, 2, sum)` (R/scavenge_compounds.R:17-22, R/run_simulation.R:249), gathered on
    the device for all cells at once.  Returns the reference's list for one cell when `cell_index`
    (0-based) is given, else the M x C matrix."""
    fmol = cells.gather(field, env_fieldVol)
    if cell_index is None:
        return fmol
    ptr, fid, dist, acc = cells.get_lists()
    sl = slice(int(ptr[cell_index]), int(ptr[cell_index + 1]))
    return {"compounds": list(env_cpds), "fmol": fmol[cell_index], "field.id": fid[sl], "field.dist": dist[sl]}


def adjust_uptake(model_react_id, model_lowbnd, cMass, accCpdFMOL, deltaTime):
    """R/adjust_uptake.R:9-33: lower bounds of the exchange reactions `EX_<compound>` the model has,
    limited by what the cell can reach in this round.  Returns (ex_react_ind 1-based, ex_react_lb) or
    (None, None) when no exchange reaction matches."""
    react_id = list(model_react_id)
    acc_mets = ["EX_" + c for c in accCpdFMOL["compounds"]]
    pos = {}
    for i, r in enumerate(react_id):
        pos.setdefault(r, i)
    rel_inds = [i for i, m in enumerate(acc_mets) if m in pos]
    if not rel_inds:
        return None, None
    model_ex_ind = np.array([pos[acc_mets[i]] for i in rel_inds], dtype=np.int64)
    fmol = np.asarray(accCpdFMOL["fmol"], dtype=np.float64)[rel_inds]
    lb1 = fmol * (1 / deltaTime) / cMass                       # fmol / (pg * hr) = mmol / (g * hr)
    lb2 = -np.asarray(model_lowbnd, dtype=np.float64)[model_ex_ind]
    return model_ex_ind + 1, -np.minimum(lb1, lb2)


def chemotaxis_shift(cells, field, pts, compound_col, cell_x, cell_y, cell_size, scavenge_dist, hill_ka, hill_coef,
                     strength, vmax, deltaTime):
    """R/run_simulation.R:260-291 for one chemotaxis compound (1-based column): the (CT.x, CT.y)
    contribution of every cell, from the moments the device takes over the cells' field lists.
    `hill_ka`, `hill_coef`, `strength`, `vmax`, `scavenge_dist` may be scalars or per-cell arrays."""
    cell_r = np.asarray(cell_size, dtype=np.float64) / 2 + scavenge_dist
    m = cells.chemotaxis_moments(field, compound_col, pts, cell_x, cell_y, np.broadcast_to(cell_r, np.shape(cell_x)))
    s0, s1, sxc, syc, sx1, sy1, n = (m[:, k] for k in range(7))
    with np.errstate(invalid="ignore", divide="ignore"):
        uniform = s0 == 0                                     # sum(f_conc) == 0 -> weights 1 (:275-276)
        centr_x = np.where(uniform, sx1 / n, sxc / s0)
        centr_y = np.where(uniform, sy1 / n, syc / s0)
        wmean = np.where(uniform, 0.0, s1 / s0)               # weighted.mean(f_conc_orig, f_conc)
        sensing = 1 / (1 + (hill_ka / wmean) ** hill_coef)    # :286
    scale = sensing * strength * vmax * 60 * deltaTime * 60
    return centr_x * scale, centr_y * scale
