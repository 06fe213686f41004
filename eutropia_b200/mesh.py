"""Field mesh + neighbour index maps (`mat.in` / `mat.out`) in the reference's format.

Host-side restatement (numpy, vectorised) of the once-per-simulation construction in
`R/growthEnvironment.R:68-153` (`initialize`) and `:164-273` (`build.DCM`).  The hot path consumes
`mat.in` / `mat.out` *as given* (SURVEY.md 8a-5); this module exists so that tests, the bench and
users without R can produce inputs in exactly that format.  It is setup code, not the per-timestep
path, and never touches the GPU.

Geometry caveat: `st_buffer` / `st_make_grid` / point-in-polygon come from sf/GEOS,
which is not vendored by the reference; the buffered polygon here is the exact offset region
{p : dist(p, polygon) <= expand} instead of GEOS's 30-segments-per-quadrant approximation, so *which*
points exist can differ at round corners.  Given the points, the neighbour map follows the
reference's rules exactly, including its one-sided snapping quirk.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import numpy as np

__all__ = [
    "FieldMesh", "universe_polygon_preset", "harbor_polygon", "make_field_points", "build_dcm",
    "make_mesh", "field_volume",
]


# ----------------------------------------------------------------------------------------------
# polygons
# ----------------------------------------------------------------------------------------------
def universe_polygon_preset(name: str) -> np.ndarray:
    """`Petri_<r>` and `Rectangle_<x>_<y>` presets (R/growthSimulation.R:165-218), plus
    `Harbor_<w>`: this repo's own synthetic non-convex polygon (stand-in for `Kiel_<w>`, whose
    vertex file is reference data and is not copied here)."""
    low = name.lower()
    if low.startswith("petri_"):
        r = float(name.split("_")[1])
        if r < 2.5:
            raise ValueError("Petri dish radius too small. (min 2.5 um)")
        t = np.linspace(0.0, 2.0 * math.pi, 100)
        poly = np.stack([np.sin(t), np.cos(t)], axis=1)
        poly[99] = poly[0]
        return poly * r
    if low.startswith("rectangle_"):
        _, sx, sy = name.split("_")
        x, y = float(sx) / 2, float(sy) / 2
        if x < 5 or y < 5:
            raise ValueError("Rectangle dimensions too small. (x,y >= 5 um)")
        return np.array([[x, y], [x, -y], [-x, -y], [-x, y]], dtype=np.float64)
    if low.startswith("harbor_"):
        return harbor_polygon(float(name.split("_")[1]))
    if low.startswith("kiel_"):
        return kiel_polygon(float(name.split("_")[1]))
    raise ValueError(f'The polygon preset "{name}" does not exist / is not yet supported.')


def kiel_csv_path() -> str | None:
    """Where the reference's `inst/extdata/kiel.csv` can be read from: $EUTROPIA_EXTDATA/kiel.csv, an
    installed R package is not looked for; /root/reference in the build container.  The file is
    reference data and is not copied into this repository: where it is absent (the GPU box), `Harbor_<w>`
    stands in for it."""
    for d in (os.environ.get("EUTROPIA_EXTDATA"), "/root/reference/inst/extdata"):
        if d and os.path.exists(os.path.join(d, "kiel.csv")):
            return os.path.join(d, "kiel.csv")
    return None


def kiel_polygon(size: float) -> np.ndarray:
    """`Kiel_<size>` (R/growthSimulation.R:202-214): the vertices of kiel.csv divided by their largest x,
    times (size / 2) / 2; sizes below 20 are refused (`x < 10` after the halving, :210-211)."""
    path = kiel_csv_path()
    if path is None:
        raise FileNotFoundError("Kiel_<L> needs the reference's inst/extdata/kiel.csv (set EUTROPIA_EXTDATA); "
                                "Harbor_<w> is this repository's own non-convex stand-in")
    kiel = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.float64)
    kiel = kiel / kiel[:, 0].max()
    x = float(size) / 2
    kiel = kiel * x / 2
    if x < 10:
        raise ValueError("Kiel latitude dimension too small (min 10 um).")
    return kiel


def harbor_polygon(width: float) -> np.ndarray:
    """A deterministic non-convex, non-star-shaped polygon (a bay with a long inlet, two
    peninsulas and a notch), scaled so that its bounding box is `width` wide."""
    v = np.array([
        [0.00, 0.10], [0.18, 0.00], [0.42, 0.04], [0.47, 0.30], [0.52, 0.62], [0.57, 0.30],
        [0.60, 0.03], [0.83, 0.00], [1.00, 0.16], [0.96, 0.40], [0.80, 0.46], [0.97, 0.58],
        [1.00, 0.82], [0.86, 1.00], [0.64, 0.93], [0.60, 0.78], [0.52, 0.74], [0.44, 0.80],
        [0.40, 0.97], [0.20, 1.00], [0.05, 0.86], [0.12, 0.66], [0.30, 0.56], [0.10, 0.48],
        [0.02, 0.30],
    ], dtype=np.float64)
    v = v - v.mean(axis=0)
    return v * float(width)


def _close(poly: np.ndarray) -> np.ndarray:
    poly = np.asarray(poly, dtype=np.float64)
    if np.any(poly[0] != poly[-1]):  # R/growthSimulation.R:126-129
        poly = np.vstack([poly, poly[:1]])
    return poly


def _inside_or_within(px: np.ndarray, py: np.ndarray, poly: np.ndarray, dist: float) -> np.ndarray:
    """Boolean mask: point is inside `poly` (boundary included) or within `dist` of its boundary."""
    x0, y0 = poly[:-1, 0], poly[:-1, 1]
    x1, y1 = poly[1:, 0], poly[1:, 1]
    out = np.zeros(px.shape[0], dtype=bool)
    chunk = max(1, 4_000_000 // max(1, x0.shape[0]))
    ex, ey = x1 - x0, y1 - y0
    el2 = ex * ex + ey * ey
    el2 = np.where(el2 == 0, 1.0, el2)
    for s in range(0, px.shape[0], chunk):
        x = px[s:s + chunk, None]
        y = py[s:s + chunk, None]
        # even-odd crossing rule
        cond = (y0[None, :] > y) != (y1[None, :] > y)
        with np.errstate(divide="ignore", invalid="ignore"):
            xint = x0[None, :] + (y - y0[None, :]) * ex[None, :] / np.where(ey == 0, 1.0, ey)[None, :]
        inside = (np.sum(cond & (x < xint), axis=1) % 2) == 1
        # distance to every segment
        t = ((x - x0[None, :]) * ex[None, :] + (y - y0[None, :]) * ey[None, :]) / el2[None, :]
        t = np.clip(t, 0.0, 1.0)
        dx = x - (x0[None, :] + t * ex[None, :])
        dy = y - (y0[None, :] + t * ey[None, :])
        d2 = np.min(dx * dx + dy * dy, axis=1)
        out[s:s + chunk] = inside | (d2 <= dist * dist)
    return out


# ----------------------------------------------------------------------------------------------
# field points  (R/growthEnvironment.R:78-118)
# ----------------------------------------------------------------------------------------------
def field_volume(field_size: float) -> float:
    """Volume of one rhombic dodecahedron (R/growthEnvironment.R:101-102)."""
    a = field_size * 3 / (2 * math.sqrt(6))
    return 16 / 9 * a ** 3 * math.sqrt(3)


def make_field_points(polygon: np.ndarray, field_size: float, field_layers: int = 3,
                      expand: float | None = None) -> np.ndarray:
    """Field-point coordinates, N x 3 (x, y, z), in the reference's layer-major order.

    Base layer: centres of a `field_size` raster over the bounding box of the polygon buffered by
    `expand` (default `field_size`), x running fastest, kept where they fall in the buffered
    polygon (:82-88).  Layer i (2..L) sits at z = field_size*(i-1); even layers are shifted by
    -field_size/2 in x and y (:104-117)."""
    poly = _close(polygon)
    fs = float(field_size)
    expand = fs if expand is None else float(expand)
    xmin, ymin = poly[:, 0].min() - expand, poly[:, 1].min() - expand
    xmax, ymax = poly[:, 0].max() + expand, poly[:, 1].max() + expand
    nx = int(math.ceil((xmax - xmin) / fs))
    ny = int(math.ceil((ymax - ymin) / fs))
    gx = xmin + (np.arange(nx, dtype=np.float64) + 0.5) * fs
    gy = ymin + (np.arange(ny, dtype=np.float64) + 0.5) * fs
    px = np.tile(gx, ny)            # x fastest (expand.grid order)
    py = np.repeat(gy, nx)
    keep = _inside_or_within(px, py, poly, expand)
    bx, by = px[keep], py[keep]
    layers = []
    for i in range(1, int(field_layers) + 1):
        lx, ly = bx.copy(), by.copy()
        if i >= 2 and i % 2 == 0:
            lx = lx - fs / 2
            ly = ly - fs / 2
        lz = np.full(bx.shape[0], fs * (i - 1), dtype=np.float64)
        layers.append(np.stack([lx, ly, lz], axis=1))
    return np.ascontiguousarray(np.vstack(layers))


# ----------------------------------------------------------------------------------------------
# neighbour maps  (R/growthEnvironment.R:164-273, :138-149)
# ----------------------------------------------------------------------------------------------
# (dx, dy, dz) in units of field_size, in the reference's melt order (:257-259); the order does not
# influence mat.in (which is re-sorted by to, from) but is kept for readability.
_OFFSETS = (
    (0.0, 1.0, 0.0), (1.0, 0.0, 0.0), (0.0, -1.0, 0.0), (-1.0, 0.0, 0.0),          # Co Cr Cu Cl
    (-0.5, 0.5, 1.0), (0.5, 0.5, 1.0), (0.5, -0.5, 1.0), (-0.5, -0.5, 1.0),        # Uol Uor Uur Uul
    (-0.5, 0.5, -1.0), (0.5, 0.5, -1.0), (0.5, -0.5, -1.0), (-0.5, -0.5, -1.0),    # Lol Lor Lur Lul
)


def _snap(q: np.ndarray, uniq: np.ndarray, tol: float) -> np.ndarray:
    """`correct_numbers` (:199-217): next-observation-carried-backward rolling join onto the sorted
    unique coordinate values (smallest value >= q; nothing beyond the last), rejected (-1) when it
    is farther than `tol` away.  Returns the index into `uniq` or -1."""
    idx = np.searchsorted(uniq, q, side="left")
    ok = idx < uniq.shape[0]
    idx_c = np.minimum(idx, uniq.shape[0] - 1)
    ok &= ~(np.abs(uniq[idx_c] - q) > tol)
    return np.where(ok, idx_c, -1)


def build_dcm(pts: np.ndarray, field_size: float) -> tuple[np.ndarray, np.ndarray]:
    """Neighbour index maps from field points.

    Returns `(mat_in, mat_out)`: `mat_in` is E x 2 int32 `(from, to)` ordered by (to, from);
    `mat_out` is N x 2 int32 `(id, n_neighbors)` with n_neighbors = out-degree + 1 (the diagonal of
    the diffusion matrix counts, :143).  1-based, like the R slots."""
    pts = np.asarray(pts, dtype=np.float64)
    n = pts.shape[0]
    fs = float(field_size)
    ux, uy, uz = np.unique(pts[:, 0]), np.unique(pts[:, 1]), np.unique(pts[:, 2])
    nux, nuy = ux.shape[0], uy.shape[0]
    ix = np.searchsorted(ux, pts[:, 0])
    iy = np.searchsorted(uy, pts[:, 1])
    iz = np.searchsorted(uz, pts[:, 2])
    key = (iz.astype(np.int64) * nuy + iy) * nux + ix
    order = np.argsort(key, kind="stable")
    skey = key[order]

    froms, tos = [], []
    ids = np.arange(1, n + 1, dtype=np.int64)
    for dx, dy, dz in _OFFSETS:
        # the shifted copy keeps the id of the point it came from (id.Co etc., :184-197) ...
        sx = _snap(pts[:, 0] + dx * fs, ux, fs / 10)
        sy = _snap(pts[:, 1] + dy * fs, uy, fs / 10)
        sz = _snap(pts[:, 2] + dz * fs, uz, fs / 10)
        ok = (sx >= 0) & (sy >= 0) & (sz >= 0)
        k = (sz[ok].astype(np.int64) * nuy + sy[ok]) * nux + sx[ok]
        # ... and is joined onto the point that owns the snapped coordinate (:239-250)
        pos = np.searchsorted(skey, k)
        pos_c = np.minimum(pos, n - 1)
        hit = (pos < n) & (skey[pos_c] == k)
        owner = order[pos_c[hit]] + 1              # id of the row in `lis` (melt: id)
        source = ids[ok][hit]                      # id.<slot> value (melt: value)
        # DCM[id, value] = 1/13, then transposed (:262-265): which() gives from = value, to = id
        froms.append(source)
        tos.append(owner)
    frm = np.concatenate(froms)
    to = np.concatenate(tos)
    # duplicates collapse (sparse assignment), order is (to, from) (:140-142)
    pair = np.unique(to * (n + 1) + frm)
    to = (pair // (n + 1)).astype(np.int32)
    frm = (pair % (n + 1)).astype(np.int32)
    mat_in = np.ascontiguousarray(np.stack([frm, to], axis=1))
    out_deg = np.bincount(frm, minlength=n + 1)[1:]
    mat_out = np.ascontiguousarray(
        np.stack([np.arange(1, n + 1, dtype=np.int32), (out_deg + 1).astype(np.int32)], axis=1))
    return mat_in, mat_out


def build_dcm_gpu(pts: np.ndarray, field_size: float, device: int = 0) -> tuple[np.ndarray, np.ndarray]:
    """`build_dcm` on the B200 (eut_build_dcm: sort / search joins, csrc/meshgen.cu); same rules, same
    integer output, bit for bit."""
    import ctypes as C

    from . import _lib
    p = np.require(np.asarray(pts, dtype=np.float64), requirements=["F", "A"])
    n = p.shape[0]
    cap = 12 * n
    buf = np.zeros((max(cap, 1), 2), dtype=np.int32, order="F")
    mat_out = np.zeros((max(n, 1), 2), dtype=np.int32, order="F")
    n_edges = C.c_int64(0)
    _lib.check(_lib.lib().eut_build_dcm(_lib.ptr(p), n, n, float(field_size), int(device), _lib.ptr(buf), cap,
                                        C.byref(n_edges), _lib.ptr(mat_out)))
    e = int(n_edges.value)
    return np.ascontiguousarray(buf[:e]), np.ascontiguousarray(mat_out[:n])


@dataclass
class FieldMesh:
    """The slots of `growthEnvironment` that the hot path reads (R/growthEnvironment.R:37-58)."""
    field_pts: np.ndarray      # N x 3 float64
    mat_in: np.ndarray         # E x 2 int32 (from, to), 1-based
    mat_out: np.ndarray        # N x 2 int32 (id, n_neighbors)
    field_size: float
    field_layers: int
    field_vol: float
    polygon: np.ndarray

    @property
    def nfields(self) -> int:
        return int(self.field_pts.shape[0])


def make_mesh(polygon, field_size: float = 1.0, field_layers: int = 3, on_device: bool = False, device: int = 0) -> FieldMesh:
    """`new("growthEnvironment", polygon.coords, field.size, field.layers)` restricted to the mesh
    slots (R/growthEnvironment.R:68-153)."""
    if isinstance(polygon, str):
        polygon = universe_polygon_preset(polygon)
    polygon = _close(polygon)
    pts = make_field_points(polygon, field_size, field_layers)
    mat_in, mat_out = build_dcm_gpu(pts, field_size, device) if on_device else build_dcm(pts, field_size)
    return FieldMesh(pts, mat_in, mat_out, float(field_size), int(field_layers),
                     field_volume(field_size), polygon)
