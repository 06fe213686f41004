"""Step (6) of a round, the part that reads the field (R/run_simulation.R:480-526):

  * `global_concentrations` -- mean of every non-constant compound over all field points (:481-486),
    reduced on the device (eut_field_column_stats);
  * `record_compounds`      -- the concentrations of the compounds of interest, rounded to 7 decimals on
    the device on their way out (eut_field_record, `round(COI.rec, digits = 7)`, :514) and saved like
    `saveRDS(data.table(...), file = <recordDir>/cpdrec_<round>_<k>.RDS, compress = T)` (:505-517).

The RDS container is written here (R serialization format 3, XDR, gzip): a list of REALSXP columns named
V1..Vk -- `@concentrations` carries no column names -- with `names`, compact `row.names` and class
c("data.table", "data.frame").  data.table's `.internal.selfref` external pointer cannot be serialised
meaningfully (R writes a nil pointer and data.table re-creates it on first use) and is left out.
`read_rds` reads the subset of the format R itself uses for such objects and for the reference's own
.RDS files; it is what the tests use to check the writer and is handy for looking at recordings without R.
`r_round` is the host mirror of the device rounding, bit for bit."""
from __future__ import annotations

import gzip
import os
import struct

import numpy as np

__all__ = ["r_round", "global_concentrations", "record_compounds", "write_rds_data_table", "read_rds"]


def _two_prod(a, b):
    """a * b = p + e exactly (Dekker / Veltkamp, no fused multiply-add: numpy has none, and the device code
    uses the same sequence so that both give the same bits)."""
    p = a * b
    c = 134217729.0                                        # 2^27 + 1
    t = c * a
    ah = t - (t - a)
    al = a - ah
    t = c * b
    bh = t - (t - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def _exact_quotient(n, d):
    """n / d as an unevaluated sum hi + lo (hi = fl(n / d), lo = the remainder's share)."""
    hi = n / d
    p, e = _two_prod(hi, d)
    lo = ((n - p) - e) / d
    return hi, lo


def r_round(x, digits: int = 7) -> np.ndarray:
    """`round(x, digits)` of base R >= 4.0.0 for digits > 0 (nmath/fround.c, restated from its published
    description -- base R is not part of the reference tree): of the two decimal neighbours
    floor(|x| * 10^d) / 10^d and ceil(|x| * 10^d) / 10^d the one closer to x AS REPRESENTED, ties to the even
    integer; |x| * 10^digits >= 1e15 and non-finite values are returned unchanged.  R compares the
    distances in long double; here they are compared through an exact quotient (double-double), which is
    that comparison without its last-bit accidents: round(0.15, 1) = 0.1, round(0.45, 1) = 0.5,
    round(0.25, 1) = 0.2 as ?Round documents.  Same IEEE operations as the device (`r_round_digits`,
    csrc/stencil.cu): bit-identical results."""
    x = np.asarray(x, dtype=np.float64)
    p10 = np.float64(10.0) ** int(digits)
    ax = np.abs(x)
    with np.errstate(invalid="ignore", over="ignore"):
        keep = ~(ax < 1e15 / p10)
        x10 = p10 * ax
        i10, c10 = np.floor(x10), np.ceil(x10)
        xd, xd_lo = _exact_quotient(i10, p10)
        xu, xu_lo = _exact_quotient(c10, p10)
        dd = (ax - xd) - xd_lo                              # distance to the lower neighbour
        du = (xu - ax) + xu_lo                              # ... and to the upper one
        even = np.fmod(i10, 2.0) == 0.0
        r = np.where((dd < du) | ((dd == du) & even), xd, xu)
    return np.where(keep, x, np.copysign(r, x))


def global_concentrations(env) -> dict:
    """`apply(concentrations[, !isConstant], 2, mean)` with the compound ids (R/run_simulation.R:481-486);
    `env` is a GrowthEnvironment."""
    _, _, mean = env.field.column_stats()
    ind_var = ~np.asarray(env.conc_isConstant, dtype=bool)
    return {"cpd.id": [c for c, v in zip(env.compounds, ind_var) if v],
            "cpd.name": [c for c, v in zip(getattr(env, "compound_names", env.compounds), ind_var) if v],
            "global_concentration": np.asarray(mean)[ind_var]}


def record_compounds(env, record, record_dir: str, sim_round: int, digits: int = 7):
    """R/run_simulation.R:498-526.  `record`: list of strings, "compounds" (all) and / or
    "compound_<id>"; returns (compound ids in file order, path of the RDS file) or (None, None) when
    nothing is to be recorded.  File name: cpdrec_<sim_round + 1>_<k>.RDS with the first unused k."""
    if not record:
        return None, None
    record = list(record)
    if not (any(r.startswith("compound_") for r in record) or "compounds" in record):
        return None, None
    coi = [r[len("compound_"):] for r in record if r.startswith("compound_")]
    coi = [c for c in coi if c in env.compounds]
    if "compounds" in record:
        coi = list(env.compounds)
    if not coi:
        return None, None
    rec_file = os.path.join(record_dir, f"cpdrec_{sim_round + 1}_0.RDS")
    k = 1
    while os.path.exists(rec_file):
        rec_file = os.path.join(record_dir, f"cpdrec_{sim_round + 1}_{k}.RDS")
        k += 1
    ind = [list(env.compounds).index(c) + 1 for c in coi]                      # match(COI, compounds)
    rec = env.field.record(ind, digits)
    write_rds_data_table(rec_file, rec)
    return coi, rec_file


# ---------------------------------------------------------------------------------------------------
# R serialization, format version 3, XDR (big endian)
# ---------------------------------------------------------------------------------------------------
_NILVALUE, _REFSXP, _ALTREP = 254, 255, 238
_SYM, _LIST, _CHAR, _LGL, _INT, _REAL, _CPLX, _STR, _VEC, _S4 = 1, 2, 9, 10, 13, 14, 15, 16, 19, 25
_HAS_ATTR, _IS_OBJ, _HAS_TAG = 1 << 9, 1 << 8, 1 << 10
_ASCII = 64 << 12
_NA_INT = -2 ** 31


def _i(v: int) -> bytes:
    return struct.pack(">i", v)


def _charsxp(s: str) -> bytes:
    b = s.encode("ascii")
    return _i(_CHAR | _ASCII) + _i(len(b)) + b


def _strsxp(v) -> bytes:
    return _i(_STR) + _i(len(v)) + b"".join(_charsxp(s) for s in v)


def _attr(name: str, value: bytes, sym_table: list) -> bytes:
    if name in sym_table:                                   # a symbol seen before is a back reference
        tag = _i((sym_table.index(name) + 1) << 8 | _REFSXP)
    else:
        sym_table.append(name)
        tag = _i(_SYM) + _charsxp(name)
    return _i(_LIST | _HAS_TAG) + tag + value


def write_rds_data_table(path: str, matrix, names=None, compress: bool = True, r_version=(4, 1, 0)) -> None:
    """`saveRDS(data.table(matrix), path, compress = T)` for a numeric matrix (column-major, one REALSXP per column)."""
    m = np.asarray(matrix, dtype=np.float64)
    if m.ndim != 2:
        raise ValueError("matrix expected")
    n, k = m.shape
    names = [f"V{j + 1}" for j in range(k)] if names is None else list(names)
    if len(names) != k:
        raise ValueError("one name per column")
    out = [b"X\n", _i(3), _i(r_version[0] << 16 | r_version[1] << 8 | r_version[2]), _i(3 << 16 | 5 << 8), _i(5), b"UTF-8"]
    out.append(_i(_VEC | _IS_OBJ | _HAS_ATTR) + _i(k))
    for j in range(k):
        out.append(_i(_REAL) + _i(n) + np.ascontiguousarray(m[:, j]).astype(">f8").tobytes())
    syms: list = []
    out.append(_attr("names", _strsxp(names), syms))
    out.append(_attr("row.names", _i(_INT) + _i(2) + _i(_NA_INT) + _i(-n), syms))      # compact form c(NA, -n)
    out.append(_attr("class", _strsxp(["data.table", "data.frame"]), syms))
    out.append(_i(_NILVALUE))
    data = b"".join(out)
    with (gzip.open(path, "wb", compresslevel=6) if compress else open(path, "wb")) as f:
        f.write(data)


class _Reader:
    def __init__(self, data: bytes):
        self.d, self.p, self.refs = data, 0, []

    def take(self, n: int) -> bytes:
        b = self.d[self.p:self.p + n]
        if len(b) != n:
            raise ValueError("truncated RDS stream")
        self.p += n
        return b

    def int(self) -> int:
        return struct.unpack(">i", self.take(4))[0]

    def length(self) -> int:
        n = self.int()
        if n == -1:                                         # long vector: two more ints
            hi, lo = self.int(), self.int()
            n = (hi << 32) + (lo & 0xffffffff)
        return n

    def item(self):
        flags = self.int()
        t = flags & 0xff
        has_attr, has_tag = bool(flags & _HAS_ATTR), bool(flags & _HAS_TAG)
        if t == _NILVALUE:
            return None
        if t == _REFSXP:
            idx = flags >> 8
            return self.refs[(idx if idx else self.int()) - 1]
        if t == _SYM:
            s = ("symbol", self.item())
            self.refs.append(s)
            return s
        if t == _CHAR:
            n = self.int()
            return None if n == -1 else self.take(n).decode("utf-8", "replace")
        if t in (_LIST, 6):                                 # pairlist / language: attr, tag, car, cdr
            out = []
            while True:
                attr = self.item() if has_attr else None
                tag = self.item() if has_tag else None
                out.append((tag[1] if tag else None, self.item()))
                flags = self.int()
                if flags & 0xff not in (_LIST, 6):
                    self.p -= 4
                    tail = self.item()
                    if tail is not None:
                        out.append((None, tail))
                    break
                has_attr, has_tag = bool(flags & _HAS_ATTR), bool(flags & _HAS_TAG)
                del attr
            return out
        if t == _ALTREP:
            info, state, attr = self.item(), self.item(), self.item()
            return {"altrep": info, "state": state, "attributes": attr}
        if t == _S4:
            return {"S4": True, "attributes": dict(self.item()) if has_attr else {}}
        if t in (4, 22, 23):                                # environment / external pointer / weak reference
            raise ValueError(f"RDS item type {t} is not supported by this reader")
        if t in (_LGL, _INT):
            n = self.length()
            v = np.frombuffer(self.take(4 * n), dtype=">i4").astype(np.int32)
        elif t == _REAL:
            n = self.length()
            v = np.frombuffer(self.take(8 * n), dtype=">f8").astype(np.float64)
        elif t == _CPLX:
            n = self.length()
            v = np.frombuffer(self.take(16 * n), dtype=">c16").astype(np.complex128)
        elif t == _STR:
            v = [self.item() for _ in range(self.length())]
        elif t in (_VEC, 20):
            v = [self.item() for _ in range(self.length())]
        elif t == 24:                                       # raw
            v = self.take(self.length())
        else:
            raise ValueError(f"RDS item type {t} is not supported by this reader")
        if has_attr:
            return {"value": v, "attributes": dict(self.item())}
        return v


def read_rds(path: str):
    """Header dict and the object of an XDR RDS file (gzip or plain).  Vectors come back as numpy arrays or
    lists, objects with attributes as {"value": ..., "attributes": {...}}, S4 objects as their slots."""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:2] == b"\x1f\x8b":
        raw = gzip.decompress(raw)
    if raw[:2] != b"X\n":
        raise ValueError("not an XDR RDS stream")
    r = _Reader(raw)
    r.p = 2
    version, writer, min_reader = r.int(), r.int(), r.int()
    header = {"version": version, "writer": (writer >> 16, (writer >> 8) & 0xff, writer & 0xff),
              "min_reader": (min_reader >> 16, (min_reader >> 8) & 0xff, min_reader & 0xff)}
    if version >= 3:
        header["encoding"] = r.take(r.int()).decode("ascii")
    obj = r.item()
    if r.p != len(raw):
        raise ValueError(f"{len(raw) - r.p} trailing bytes after the object")
    return header, obj
