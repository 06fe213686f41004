"""Deterministic synthetic inputs for tests and the bench (SURVEY.md 8d "Value distributions").

Input generation only: nothing here is on the compute path.
"""
from __future__ import annotations

import math

import numpy as np

from . import mesh as _mesh

__all__ = ["square_side_for", "harbor_width_for", "synthetic_concentrations", "place_cells",
           "synthetic_exchanges", "DEFAULT_CELL_DIAMETER", "DEFAULT_SCAVENGE_DIST", "config_mesh"]

# add_organism defaults (R/growthSimulation.R:332, :337)
DEFAULT_CELL_DIAMETER = (3 * 1 / (4 * math.pi)) ** (1 / 3) * 2
DEFAULT_SCAVENGE_DIST = DEFAULT_CELL_DIAMETER * 2.5


def square_side_for(n_points: int, layers: int = 3, field_size: float = 1.0) -> int:
    """Side s of `Rectangle_<s>_<s>` so that the mesh has about n_points (base = (s+2)^2 at fs 1)."""
    base = n_points / layers
    return max(10, int(round(math.sqrt(base) * field_size - 2 * field_size)))


def harbor_width_for(n_points: int, layers: int = 3, field_size: float = 1.0) -> float:
    """Width of `Harbor_<w>` whose mesh has about n_points."""
    poly = _mesh.harbor_polygon(1.0)
    x, y = poly[:, 0], poly[:, 1]
    area = 0.5 * abs(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))
    return math.sqrt(n_points / layers * field_size ** 2 / area)


def _kiel_mesh(size: int):
    """`Kiel_<size>`, 3 layers: from the reference's inst/extdata/kiel.csv where that file exists (the build
    container, mesh.kiel_csv_path), else from field points written there by tools/cache_kiel_points.py into
    the git-ignored directory $EUT_MESH_CACHE (default <repo>/mesh_cache) -- derived points, not the
    reference's file, which is never copied -- so that the configuration can be measured on the GPU box."""
    import os
    if _mesh.kiel_csv_path() is not None:
        return _mesh.make_mesh(f"Kiel_{size}", 1.0, 3)
    cache = os.environ.get("EUT_MESH_CACHE") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mesh_cache")
    path = os.path.join(cache, f"kiel_{size}_pts.npz")
    if not os.path.exists(path):
        raise FileNotFoundError(f"Kiel_{size}: neither the reference's kiel.csv nor {path} is here; use harbor1m")
    z = np.load(path)
    pts = np.ascontiguousarray(z["pts"], dtype=np.float64)
    mat_in, mat_out = _mesh.build_dcm(pts, 1.0)
    return _mesh.FieldMesh(pts, mat_in, mat_out, 1.0, 3, _mesh.field_volume(1.0), z["polygon"])


def config_mesh(name: str, on_device: bool = False, device: int = 0):
    """The BASELINE.json configs as meshes (`on_device`: neighbour maps by eut_build_dcm)."""
    if on_device:
        import functools
        real = _mesh.make_mesh
        _mesh.make_mesh = functools.partial(real, on_device=True, device=device)
        try:
            return config_mesh(name)
        finally:
            _mesh.make_mesh = real
    if name == "square100k":      # config 2: 183 x 183 x 3 = 100 467 points
        return _mesh.make_mesh("Rectangle_181_181", 1.0, 3)
    if name == "harbor1m":        # config 3/4: non-convex polygon, ~1e6 points, 3 layers
        return _mesh.make_mesh(_mesh.harbor_polygon(harbor_width_for(1_000_000)), 1.0, 3)
    if name == "kiel1m":          # config 3 on the reference's own polygon, `Kiel_1712` (~1e6 points at 3 layers)
        return _kiel_mesh(1712)
    if name == "petri80":         # config 1 (vignette): Petri_80, 6 layers
        return _mesh.make_mesh("Petri_80", 1.0, 6)
    if name == "square16m":       # config 5
        s = square_side_for(16_000_000)
        return _mesh.make_mesh(f"Rectangle_{s}_{s}", 1.0, 3)
    if name.startswith("squareL"):      # squareL<layers>:<points>
        layers = int(name[7:name.index(":")])
        s = square_side_for(int(name.split(":")[1]), layers=layers)
        return _mesh.make_mesh(f"Rectangle_{s}_{s}", 1.0, layers)
    if name.startswith("square:"):
        s = square_side_for(int(name.split(":")[1]))
        return _mesh.make_mesh(f"Rectangle_{s}_{s}", 1.0, 3)
    if name.startswith("harbor:"):
        return _mesh.make_mesh(_mesh.harbor_polygon(harbor_width_for(int(name.split(":")[1]))), 1.0, 3)
    raise ValueError(name)


def synthetic_concentrations(pts: np.ndarray, n_cols: int, seed: int, out=None) -> np.ndarray:
    """mM, non-negative: uniform[0, 25] base + sparse Gaussian blobs, a block of exact zeros, every
    8th column near 1000 (medium.csv scale).  Column-major N x n_cols."""
    n = pts.shape[0]
    rng = np.random.default_rng(seed)
    if out is None:
        out = np.empty((n, n_cols), dtype=np.float64, order="F")
    x, y = pts[:, 0], pts[:, 1]
    span = max(float(x.max() - x.min()), 1.0)
    for c in range(n_cols):
        col = rng.uniform(0.0, 25.0, n)
        for _ in range(3):
            k = rng.integers(0, n)
            s = span * rng.uniform(0.01, 0.05)
            col += 40.0 * np.exp(-((x - x[k]) ** 2 + (y - y[k]) ** 2) / (2 * s * s))
        if c % 8 == 7:
            col += 1000.0
        z0 = rng.integers(0, max(1, n - n // 50))
        col[z0:z0 + n // 50] = 0.0
        out[:, c] = col
    return out


def place_cells(mesh, n_cells: int, seed: int):
    """Cell centres uniform over the universe polygon (rejection sampling), default geometry."""
    rng = np.random.default_rng(seed)
    poly = mesh.polygon
    lo, hi = poly.min(axis=0), poly.max(axis=0)
    xs, ys = [], []
    need = n_cells
    while need > 0:
        cx = rng.uniform(lo[0], hi[0], 2 * need + 16)
        cy = rng.uniform(lo[1], hi[1], 2 * need + 16)
        ok = _mesh._inside_or_within(cx, cy, poly, 0.0)
        xs.append(cx[ok][:need]); ys.append(cy[ok][:need])
        need -= xs[-1].shape[0]
    cx, cy = np.concatenate(xs), np.concatenate(ys)
    size = np.full(n_cells, DEFAULT_CELL_DIAMETER)
    scv = np.full(n_cells, DEFAULT_SCAVENGE_DIST)
    return cx, cy, size, scv


def synthetic_exchanges(n_cells: int, n_cols: int, per_cell: int, seed: int):
    """Per-cell exchange lists: `per_cell` distinct compound columns (1-based) with N(0,1) fmol
    fluxes (about half uptake, half secretion).  Returns (ex_ptr int64, ex_cpd int32, ex_flx)."""
    rng = np.random.default_rng(seed)
    per_cell = min(per_cell, n_cols)
    ex_ptr = np.arange(0, (n_cells + 1) * per_cell, per_cell, dtype=np.int64)
    # distinct columns per cell: random start + stride walk (cheap, no per-cell shuffles)
    start = rng.integers(0, n_cols, n_cells)
    step = np.ones(n_cells, dtype=np.int64)
    offs = np.arange(per_cell, dtype=np.int64)
    cpd = ((start[:, None] + offs[None, :] * step[:, None]) % n_cols + 1).astype(np.int32)
    flx = rng.normal(0.0, 1.0, (n_cells, per_cell))
    flx[flx == 0] = 1.0
    return ex_ptr, np.ascontiguousarray(cpd.ravel()), np.ascontiguousarray(flx.ravel())
