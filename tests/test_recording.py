"""Step (6) of a round, the part that reads the field (R/run_simulation.R:480-526): global concentrations,
rounding to 7 decimals, the cpdrec_<round>_<k>.RDS files.  base R (`round`, `saveRDS`) is third-party code
outside the reference tree: the rounding rule is pinned by the answers ?Round documents for R >= 4.0.0, the
container by the reference's own R-written .RDS files (read back here where /root/reference exists)."""
import gzip
import os
import struct
import types

import numpy as np
import pytest

import eutropia_b200 as eb
from eutropia_b200 import recording


def test_r_round_documented_answers():
    # ?Round, "round to even on the represented number": 0.15 -> 0.1, 0.25 -> 0.2, 0.35 -> 0.3, 0.45 -> 0.5, ...
    assert recording.r_round([0.15, 0.25, 0.35, 0.45, 0.55, 0.65], 1).tolist() == [0.1, 0.2, 0.3, 0.5, 0.6, 0.7]
    assert recording.r_round([2.675, 1.005, 0.285], 2).tolist() == [2.67, 1.0, 0.28]          # all represented below the tie
    x = np.array([0.0, -0.0, 5e-8, 1.5e-7, 2.5e-7, -1234.56789012345, 25.0, 1e-12, 0.12345675, 1e9 + 0.123456789,
                  np.nan, np.inf, -np.inf])
    r = recording.r_round(x, 7)
    assert r[:10].tolist() == [0.0, -0.0, 0.0, 1e-7, 2e-7, -1234.5678901, 25.0, 0.0, 0.1234568, 1e9 + 0.123456789]
    assert np.signbit(r[1]) and np.isnan(r[10]) and r[11] == np.inf and r[12] == -np.inf
    # idempotent, and never farther than half a unit of the last place kept
    rng = np.random.default_rng(0)
    v = rng.uniform(0, 30, 20000) * 10.0 ** rng.integers(-9, 3, 20000)
    rv = recording.r_round(v, 7)
    assert np.array_equal(recording.r_round(rv, 7), rv)
    assert np.all(np.abs(rv - v) <= 0.5e-7 * (1 + 1e-9))


def test_rds_writer_round_trip_and_header(tmp_path):
    rng = np.random.default_rng(1)
    m = recording.r_round(rng.uniform(0, 25, (37, 3)), 7)
    path = str(tmp_path / "cpdrec_1_0.RDS")
    recording.write_rds_data_table(path, m)
    raw = open(path, "rb").read()
    assert raw[:2] == b"\x1f\x8b"                                      # saveRDS(compress = T): gzip
    data = gzip.decompress(raw)
    # "X\n", format 3, writer R 4.1.0, readable from R 3.5.0, native encoding "UTF-8": the bytes R 4.1.0 wrote at
    # the head of the reference's inst/extdata/eure.RDS
    assert data[:23] == b"X\n" + struct.pack(">iii", 3, 0x00040100, 0x00030500) + struct.pack(">i", 5) + b"UTF-8"
    assert struct.unpack(">ii", data[23:31]) == (19 | 0x100 | 0x200, 3)     # VECSXP, object, attributes; 3 columns
    header, obj = recording.read_rds(path)
    assert header == {"version": 3, "writer": (4, 1, 0), "min_reader": (3, 5, 0), "encoding": "UTF-8"}
    assert obj["attributes"]["names"] == ["V1", "V2", "V3"]                 # data.table(matrix without colnames)
    assert obj["attributes"]["class"] == ["data.table", "data.frame"]
    assert obj["attributes"]["row.names"].tolist() == [-2 ** 31, -37]       # compact row names c(NA, -n)
    assert all(np.array_equal(obj["value"][j], m[:, j]) for j in range(3))
    recording.write_rds_data_table(path, m[:, :1], names=["glc"], compress=False)
    _, obj = recording.read_rds(path)
    assert obj["attributes"]["names"] == ["glc"] and np.array_equal(obj["value"][0], m[:, 0])


@pytest.mark.skipif(not os.path.exists("/root/reference/inst/extdata/eure.RDS"), reason="the reference's R-written files are not on this machine")
def test_reader_understands_files_written_by_r():
    """The reader is only worth something as the writer's checker if it reads what R writes: the three
    model files of the reference (S4 objects with data frames, strings, sparse matrices; R 3.6.1 and 4.1.0)
    parse to their last byte, and R's own data frames carry the attribute layout the writer emits."""
    for name, writer in (("eure", (4, 1, 0)), ("bilo", (4, 1, 0)), ("zymo", (3, 6, 1))):
        header, obj = recording.read_rds(f"/root/reference/inst/extdata/{name}.RDS")
        assert header["version"] == 3 and header["writer"] == writer and header["encoding"] == "UTF-8"
        df = obj["attributes"]["react_attr"]
        n = len(df["value"][0])
        assert df["attributes"]["class"] == ["data.frame"] and df["attributes"]["row.names"].tolist() == [-2 ** 31, -n]
        assert len(df["attributes"]["names"]) == len(df["value"])


class _StubField:
    def __init__(self, conc):
        self.conc = conc

    def record(self, cols, digits=7):
        return recording.r_round(self.conc[:, np.asarray(cols) - 1], digits)

    def column_stats(self):
        return self.conc.min(0), self.conc.max(0), self.conc.mean(0)


def test_record_compounds_selection_and_file_names(tmp_path):
    conc = np.random.default_rng(2).uniform(0, 25, (11, 4))
    env = types.SimpleNamespace(compounds=["glc", "o2", "ac", "h2o"], conc_isConstant=[False, True, False, False],
                                field=_StubField(conc))
    d = str(tmp_path)
    assert recording.record_compounds(env, None, d, 0) == (None, None)
    assert recording.record_compounds(env, ["cells"], d, 0) == (None, None)
    assert recording.record_compounds(env, ["compound_xyz"], d, 0) == (None, None)          # not in the environment
    coi, f0 = recording.record_compounds(env, ["compound_ac", "compound_xyz", "compound_glc"], d, 4)
    assert coi == ["ac", "glc"] and os.path.basename(f0) == "cpdrec_5_0.RDS"               # :506, record order
    _, obj = recording.read_rds(f0)
    assert np.array_equal(obj["value"][0], recording.r_round(conc[:, 2], 7)) and np.array_equal(obj["value"][1], recording.r_round(conc[:, 0], 7))
    coi, f1 = recording.record_compounds(env, ["compounds", "compound_ac"], d, 4)           # "compounds": all of them (:503-504)
    assert coi == env.compounds and os.path.basename(f1) == "cpdrec_5_1.RDS"               # first unused k (:507-511)
    g = recording.global_concentrations(env)
    assert g["cpd.id"] == ["glc", "ac", "h2o"] and np.array_equal(g["global_concentration"], conc.mean(0)[[0, 2, 3]])


@pytest.mark.gpu
def test_device_recording_matches_host_mirror(tmp_path):
    m = eb.mesh.make_mesh(eb.mesh.harbor_polygon(60.0), 1.0, 3)
    rng = np.random.default_rng(3)
    conc = np.asfortranarray(rng.uniform(0, 25, (m.nfields, 5)) * 10.0 ** rng.integers(-9, 2, (m.nfields, 5)))
    conc[::7, 1] = 0.0
    conc[1::7, 1] = 2.5e-7                                            # a tie at the 7th decimal, as represented
    conc[:, 3] = 1000.0
    env = eb.GrowthEnvironment(m, ["a", "b", "c", "d", "e"], conc, [75.0] * 5, conc_isConstant=[0, 0, 0, 1, 0])
    got = env.field.record([2, 5, 1], 7)
    assert np.array_equal(got, recording.r_round(conc[:, [1, 4, 0]], 7))              # same IEEE operations: same bits
    assert np.array_equal(env.field.download(), conc)                                  # the field itself is not rounded
    coi, path = recording.record_compounds(env, ["compounds"], str(tmp_path), 0)
    _, obj = recording.read_rds(path)
    assert coi == ["a", "b", "c", "d", "e"] and all(np.array_equal(obj["value"][j], recording.r_round(conc[:, j], 7)) for j in range(5))
    g = recording.global_concentrations(env)
    assert g["cpd.id"] == ["a", "b", "c", "e"]
    assert np.all(np.abs(g["global_concentration"] - conc.mean(0)[[0, 1, 2, 4]]) <= 1e-12 * conc.mean(0)[[0, 1, 2, 4]])
    with pytest.raises(eb.EutropiaError):
        env.field.record([1], 0)
