"""Known-answer fixture for the R-language steps of the hot path (cell -> field lists, claim rescale,
gather, scatter + clamp, coordinate snapping), written FROM THE R SEMANTICS, statement by statement, in
plain Python: lists, floats, `math` -- no numpy, nothing from oracle/ or eutropia_b200/.  The image has
no R, so this is the closest thing to a reference run that can exist here; every block cites the R lines
it transcribes.  `tests/golden/make_r_golden.R` writes the same JSON from a real R session with Eutropia,
sf and data.table installed -- a maintainer who has R replaces `r_semantics.json` with its output and the
tests of tests/test_r_semantics.py then pin the oracle and the device code to R itself.

Assumptions nobody could check here are listed under "assumptions" in the JSON (GEOS distance is planar;
`sum()` accumulates in long double -- the inputs are chosen so that it does not matter).

    python tests/golden/make_r_semantics.py        # rewrites tests/golden/r_semantics.json
"""
import json
import math
import os

here = os.path.dirname(os.path.abspath(__file__))

# ---------------------------------------------------------------------------------------------------
# field points: two layers, the second shifted by -fs/2 in x and y (R/growthEnvironment.R:104-117),
# listed in a deliberately scrambled order inside each layer: the field id is the row number
# (R/run_simulation.R:167 `field := 1:.N`), the ORDER of a cell's list is the data.table key (z, x, y)
# (:168), not the id.
# ---------------------------------------------------------------------------------------------------
layer1 = [(x, y, 0.0) for y in (2.0, 0.0, 1.0) for x in (3.0, 1.0, 0.0, 2.0)]
layer2 = [(x, y, 1.0) for x in (1.5, -0.5, 2.5, 0.5) for y in (0.5, 1.5, -0.5)]
pts = layer1 + layer2
N = len(pts)
field_vol = 16.0 / 9.0 * (1.0 * 3.0 / (2.0 * math.sqrt(6.0))) ** 3 * math.sqrt(3.0)   # R/growthEnvironment.R:101-102, fs = 1

cells = [  # x, y, size, scavenge distance
    (1.0, 1.0, 1.0, 1.0),
    (2.0, 1.0, 1.0, 1.0),
]


def cell_field_env(cx, cy, size, scv):
    """R/run_simulation.R:225-227 (cube filter on the keyed table) + :591-624 (get_cellFieldEnv_ex)."""
    R = size / 2 + scv
    grid = sorted(range(N), key=lambda i: (pts[i][2], pts[i][0], pts[i][1]))           # setkeyv(gridDT, c("z","x","y"))  :168
    cube = [i for i in grid if abs(pts[i][0] - cx) <= R and abs(pts[i][1] - cy) <= R and abs(pts[i][2] - 0) <= R]   # :225-227
    out = []
    for i in cube:
        dx, dy = pts[i][0] - cx, pts[i][1] - cy
        d = math.sqrt(dx * dx + dy * dy)          # st_distance(XYZ points, st_point(c(x, y, 0))): GEOS, planar (assumption)  :609
        if d <= R:                                # :612
            acc = -1 / scv * (d - R)              # :618   ((-1) / scv) * (d - R) by R's precedence
            if acc > 1:
                acc = 1.0                         # :619
            out.append((i + 1, d, acc))
    return out


lists = [cell_field_env(*c) for c in cells]

# ---- claim rescale, R/run_simulation.R:236-241 -------------------------------------------------------
table = [(ci, fid, d, acc) for ci, lst in enumerate(lists) for (fid, d, acc) in lst]     # rbindlist(..., idcol = T)
claim = {}
for _, fid, _, acc in table:
    claim[fid] = claim.get(fid, 0.0) + acc                 # field_claim_sum := sum(acc.prop), by = field.id (table order)
tmp_sum = {}
for _, fid, _, acc in table:
    if claim[fid] > 1:
        tmp_sum[fid] = tmp_sum.get(fid, 0.0) + acc ** math.exp(1)
acc_after = []
for _, fid, _, acc in table:
    if claim[fid] > 1:
        t = acc ** math.exp(1)                             # acc.prop.tmp := acc.prop^exp(1)
        t = t / tmp_sum[fid] * claim[fid]                  # acc.prop.tmp / sum(acc.prop.tmp) * field_claim_sum
        acc_after.append(t / claim[fid])                   # acc.prop := acc.prop.tmp / field_claim_sum
    else:
        acc_after.append(acc)

# ---- concentrations ----------------------------------------------------------------------------------
# compound 1: dyadic values; compound 2: zero everywhere (colSums == 0 -> 0/0); compound 3: constant;
# compound 4: small values that the uptake drives below 1e-12 on some fields
conc = [[0.0] * 4 for _ in range(N)]
for i in range(N):
    conc[i][0] = 0.5 + 0.25 * (i % 5)
    conc[i][1] = 0.0
    conc[i][2] = 7.0
    conc[i][3] = 2.0 ** -30 * (1 + (i % 3))
untouched = pts.index((-0.5, -0.5, 1.0))                   # in neither cell's list
conc[untouched][0] = 5e-13                                 # < 1e-12 but never touched: must stay  (:319 acts on I[,1:2] only)
is_constant = [False, False, True, False]

# ---- gather: scavenge_compounds, R/scavenge_compounds.R:17-22 ------------------------------------------
ptr = [0]
for lst in lists:
    ptr.append(ptr[-1] + len(lst))
fmol_per_field = []          # per table row, per compound
for k, (ci, fid, d, _) in enumerate(table):
    a = acc_after[k]
    fmol_per_field.append([conc[fid - 1][c] / 1000 * field_vol * a for c in range(4)])    # ((conc / 1000) * fieldVol) * acc.prop
fmol = [[math.fsum(fmol_per_field[k][c] for k in range(ptr[ci], ptr[ci + 1])) for c in range(4)] for ci in range(len(cells))]

# ---- adjust_uptake, R/adjust_uptake.R:27-32 ------------------------------------------------------------
delta_time, c_mass = 1.0 / 6.0, [0.5, 0.75]
lowbnd = [-10.0, -1000.0, -1e-9, 0.0]                      # of EX_cpd1 .. EX_cpd4
ex_react_lb = [[-min(fmol[ci][c] * (1 / delta_time) / c_mass[ci], -lowbnd[c]) for c in range(4)] for ci in range(len(cells))]

# ---- scatter: R/run_simulation.R:682-740 (per cell) and :309-319 (aggregate, apply, clamp) --------------
# ex.flx in fmol per round (flux * cMass * deltaTime already applied, :687); compound ids are 1-based columns
exchanges = [
    [(1, -3.0e-4), (2, +2.0e-3), (3, +1.0e-3), (4, -1.0)],   # cell 1: uptake 1, production 2, production of a constant compound, huge uptake of 4
    [(1, -5.0e-4), (2, -1.0e-3), (4, +4.0e-6)],              # cell 2: uptake 1, uptake of compound 2 (colSums == 0 -> NaN), production of 4
]
vol_l = field_vol / 1e15


def rdiv(a, b):
    """R's `/` on doubles: x / 0 is +-Inf, 0 / 0 is NaN."""
    if b != 0:
        return a / b
    return float("nan") if a == 0 or a != a else math.copysign(float("inf"), a)


triplets = []                                               # (field.id, concMatInd, concChange) in rbindlist order
for ci in range(len(cells)):
    rows = range(ptr[ci], ptr[ci + 1])
    col_sums = [math.fsum(fmol_per_field[k][c] for k in rows) for c in range(4)]          # colSums(accmat)  :696
    up = [(c, ex) for c, ex in exchanges[ci] if ex < 0]
    pd = [(c, ex) for c, ex in exchanges[ci] if ex > 0]
    for c, ex in up:                                                                        # I.up: column by column  :707-714
        for k in rows:
            am = rdiv(fmol_per_field[k][c - 1], col_sums[c - 1])                            # accmat / colSums  :696
            triplets.append((table[k][1], c, ((am * ex) / 1e12) / vol_l))
    if pd:
        dmax = max(table[k][2] for k in rows)
        frac = [dmax - table[k][2] for k in rows]                                           # :729
        fsum = math.fsum(frac)
        frac = [f / fsum for f in frac]                                                     # :730
        for c, ex in pd:
            e1 = (ex / 1e12) / vol_l                                                        # :732
            for j, k in enumerate(rows):
                triplets.append((table[k][1], c, frac[j] * e1))                             # field.frac %*% t(ex.pro)  :734
groups = {}
order = []
for fid, c, d in triplets:                                   # I[, sum(concChange), by = c("field.id", "concMatInd")]  :310
    if (fid, c) not in groups:
        groups[(fid, c)] = 0.0
        order.append((fid, c))
    groups[(fid, c)] = groups[(fid, c)] + d
after = [row[:] for row in conc]
touched = []
for fid, c in order:
    if is_constant[c - 1]:                                   # :311-312
        continue
    d = groups[(fid, c)]
    if d != d:
        d = 0.0                                              # ifelse(is.na(I[,3]), 0, I[,3])  :315  (is.na(NaN) is TRUE)
    v = after[fid - 1][c - 1] + d                            # :318
    if v < 1e-12:
        v = 0.0                                              # :319
    after[fid - 1][c - 1] = v
    touched.append([fid, c])

# ---- correct_numbers, R/growthEnvironment.R:199-217 ------------------------------------------------------
# all_x[J(qn), roll = -Inf]: next observation carried backward = the smallest table value >= qn (NA beyond
# the last), rejected when farther than field.size / 10
all_x = [0.0, 0.5, 1.0, 1.5000000000000002, 2.0]
queries = [0.5, 0.49999999999999994, 0.96, 0.89, 1.5, 1.5000000000000004, 2.0000000000000004, -0.05, -0.2]


def snap(q):
    cand = [v for v in all_x if v >= q]
    if not cand:
        return None
    v = cand[0]
    return None if abs(v - q) > 1.0 / 10 else v


snapped = [snap(q) for q in queries]

out = {
    "about": "Known answers for the R-language steps, transcribed statement by statement from the reference "
             "(see make_r_semantics.py for the R line of every step); replace with the output of make_r_golden.R "
             "to pin against R itself.",
    "generated_by": "tests/golden/make_r_semantics.py (plain Python, no R available in this image)",
    "assumptions": ["sf::st_distance on XYZ points is GEOS' planar distance (Z ignored)",
                    "inputs chosen so that the accumulator width of sum()/colSums() (long double in R) cannot matter beyond 1e-15 relative"],
    "field_pts": [list(p) for p in pts], "field_vol": field_vol,
    "cells": [list(c) for c in cells],
    "cell_ptr": ptr,
    "field_id": [t[1] for t in table], "field_dist": [t[2] for t in table],
    "acc_raw": [t[3] for t in table], "acc_prop": acc_after,
    "claimed_fields": sorted(f for f, s in claim.items() if s > 1),
    "conc": conc, "is_constant": is_constant,
    "fmol": fmol,
    "adjust_uptake": {"delta_time": delta_time, "c_mass": c_mass, "lowbnd": lowbnd, "ex_react_lb": ex_react_lb},
    "exchanges": exchanges, "after": after, "touched": touched, "untouched_small_entry": [untouched + 1, 1],
    "snap": {"all": all_x, "queries": queries, "snapped": snapped, "field_size": 1.0},
}
with open(os.path.join(here, "r_semantics.json"), "w") as f:
    json.dump(out, f, indent=1)
print("r_semantics.json:", N, "points,", len(table), "entries,", len(order), "touched (field, compound) groups,",
      len(out["claimed_fields"]), "oversubscribed fields")
