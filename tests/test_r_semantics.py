"""The R-language steps of the hot path against tests/golden/r_semantics.json: known answers transcribed
statement by statement from the reference's R code (tests/golden/make_r_semantics.py; every step cites
its R line) or -- once a maintainer with R has run tests/golden/make_r_golden.R -- written by R itself.
The CPU tests hold the oracle (oracle/cellfield_oracle.c) and the host mirrors to it, the GPU tests the
device code through the C ABI.

Quirks pinned here: list order = data.table key (z, x, y), not the field id (R/run_simulation.R:168);
`<=` on the cube and on the distance (:225-227, :612); planar distance (assumption, see the JSON);
acc.prop capped at 1 (:619); claim rescale only where the claim sum is > 1, acc^e / sum(acc^e) (:236-241);
gather in the order ((c / 1000) * V) * acc (R/scavenge_compounds.R:17-18); by-(field, compound) sums over
the concatenated per-cell tables, a NaN (0 / 0 from an all-zero column) poisons its whole group and is
then counted as 0 (:310, :315); constant compounds are dropped (:311-312); `< 1e-12 -> 0` on the touched
entries only (:319); next-observation-carried-backward snapping with the fs / 10 rejection
(R/growthEnvironment.R:199-217)."""
import json
import os

import numpy as np
import pytest

import oracle
import eutropia_b200 as eb
from eutropia_b200 import exchange, mesh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "r_semantics.json")


@pytest.fixture(scope="module")
def rs():
    with open(GOLD) as f:
        d = json.load(f)
    d["pts"] = np.array(d["field_pts"], dtype=np.float64)
    d["conc_m"] = np.asfortranarray(np.array(d["conc"], dtype=np.float64))
    d["after_m"] = np.array(d["after"], dtype=np.float64)
    cl = np.array(d["cells"], dtype=np.float64)
    d["cx"], d["cy"], d["size"], d["scv"] = (np.ascontiguousarray(cl[:, k]) for k in range(4))
    ex_ptr, ex_cpd, ex_flx = [0], [], []
    for lst in d["exchanges"]:
        for c, f in lst:
            ex_cpd.append(c); ex_flx.append(f)
        ex_ptr.append(len(ex_cpd))
    d["ex"] = (np.array(ex_ptr, dtype=np.int64), np.array(ex_cpd, dtype=np.int32), np.array(ex_flx, dtype=np.float64))
    return d


def ulps(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.spacing(np.abs(b)), 1e-300)) if a.size else 0.0


def check_lists(d, ptr, fid, fd, acc_raw, acc_prop):
    assert ptr.tolist() == d["cell_ptr"]
    assert fid.tolist() == d["field_id"]                        # index map: exact, key order (z, x, y)
    assert np.array_equal(fd, np.array(d["field_dist"]))        # sqrt(dx*dx + dy*dy): same IEEE operations
    assert np.array_equal(acc_raw, np.array(d["acc_raw"]))
    claimed = np.isin(fid, d["claimed_fields"])
    assert claimed.any() and not claimed.all()
    assert np.array_equal(acc_prop[~claimed], np.array(d["acc_raw"])[~claimed])      # untouched where the claim sum is <= 1
    assert ulps(acc_prop[claimed], np.array(d["acc_prop"])[claimed]) <= 4             # pow() is libm's: a few ulp


def check_after(d, got):
    want = d["after_m"]
    assert np.all(np.abs(got - want) <= 1e-13 * np.abs(want))
    assert np.array_equal(got == 0.0, want == 0.0)                                    # the clamp hits the same entries
    f, c = (int(v) for v in d["untouched_small_entry"])
    assert got[f - 1, c - 1] == d["conc"][f - 1][c - 1] == 5e-13                     # below 1e-12 but untouched: stays
    assert np.array_equal(got[:, 2], d["conc_m"][:, 2])                               # constant compound
    touched = np.zeros(want.shape, dtype=bool)
    for f, c in d["touched"]:
        touched[int(f) - 1, int(c) - 1] = True
    assert np.array_equal(got[~touched], d["conc_m"][~touched])


def test_oracle_matches_r_known_answers(rs):
    d = rs
    ptr, fid, fd, acc = oracle.cellmap(d["pts"], d["cx"], d["cy"], d["size"], d["scv"])
    acc2 = oracle.claim(fid, acc, d["pts"].shape[0])
    check_lists(d, ptr, fid, fd, acc, acc2)
    fmol = oracle.gather(d["conc_m"], ptr, fid, acc2, d["field_vol"])
    assert np.all(np.abs(fmol - np.array(d["fmol"])) <= 1e-14 * np.abs(np.array(d["fmol"])))
    after = oracle.scatter(d["conc_m"], d["is_constant"], ptr, fid, fd, acc2, d["field_vol"], *d["ex"])
    check_after(d, after)


def test_host_mirrors_match_r_known_answers(rs):
    d = rs
    # correct_numbers: NOCB rolling join + rejection beyond fs / 10
    s = d["snap"]
    uniq = np.array(s["all"])
    idx = mesh._snap(np.array(s["queries"]), uniq, s["field_size"] / 10)
    got = [None if i < 0 else float(uniq[i]) for i in idx]
    assert got == s["snapped"]
    # adjust_uptake: -min(fmol / dt / cMass, -lowbnd)
    au = d["adjust_uptake"]
    react = ["EX_cpd%d" % (k + 1) for k in range(4)]
    for ci in range(len(d["cells"])):
        ind, lb = exchange.adjust_uptake(react, au["lowbnd"], au["c_mass"][ci],
                                         {"compounds": ["cpd%d" % (k + 1) for k in range(4)], "fmol": d["fmol"][ci]},
                                         au["delta_time"])
        assert ind.tolist() == [1, 2, 3, 4]
        assert np.array_equal(lb, np.array(au["ex_react_lb"][ci]))


@pytest.mark.gpu
def test_device_matches_r_known_answers(rs):
    d = rs
    mat_in, mat_out = mesh.build_dcm(d["pts"], 1.0)
    plan = eb.Plan(mat_in, mat_out)
    cells = eb.Cells(plan)
    cells.build_lists(d["pts"], d["cx"], d["cy"], d["size"], d["scv"])
    ptr, fid, fd, acc = cells.get_lists()
    acc = acc.copy()
    cells.claim_rescale()
    _, _, _, acc2 = cells.get_lists()
    check_lists(d, ptr, fid, fd, acc, acc2)
    fld = eb.Field(plan, 4)
    fld.upload(d["conc_m"])
    fmol = cells.gather(fld, d["field_vol"])
    assert np.all(np.abs(fmol - np.array(d["fmol"])) <= 1e-12 * np.abs(np.array(d["fmol"])))
    cells.scatter(fld, d["field_vol"], d["is_constant"], *d["ex"])
    check_after(d, fld.download())
