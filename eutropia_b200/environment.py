"""Host-side mirror of the `growthEnvironment` slots the hot path reads and of its dispatcher
`diffuse_compoundsPar` (R/growthEnvironment.R:37-58, :367-437), with the field resident on a B200.

The dispatcher logic is R code in the reference and "stays in R" (SURVEY.md 8a-4); it is mirrored
here so that the whole round -- column selection, substep counts, D = Inf rule, diffusion -- can be
driven and tested without R, and so that a resident session never has to pull the field to the host
just to decide which columns diffuse.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as _dc_field

import numpy as np

from . import rcpp_api
from .mesh import FieldMesh
from .session import Cells, Field, Plan

__all__ = ["GrowthEnvironment", "Exoenzyme", "diffusion_niter", "indDiff_worker", "chunk_codes"]


@dataclass
class Exoenzyme:
    """The slots of the reference's `Exoenzyme` class the hot path reads (R/run_simulation.R:343-371,
    R/growthEnvironment.R:406-434): `mets[0]` is the substrate, `stoich[i]` multiplies dS."""
    id: str
    mets: list
    stoich: list
    Kcat: float
    Km: float
    D: float = 10.0
    lambda_: float = 0.0
    name: str = ""
    extra: dict = _dc_field(default_factory=dict)


def chunk_codes(n: int, n_chunks: int) -> np.ndarray:
    """`cut(seq_len(n), n_chunks, labels = FALSE)` (base R cut.default with a break count), as
    `diffuse_compounds` uses it to split the diffusing columns (R/growthEnvironment.R:299-301)."""
    x = np.arange(1, n + 1, dtype=np.float64)
    nb = int(n_chunks) + 1
    lo, hi = 1.0, float(n)
    dx = hi - lo
    if dx == 0:
        dx = abs(lo)
        breaks = np.linspace(lo - dx / 1000, hi + dx / 1000, nb)
    else:
        breaks = np.linspace(lo, hi, nb)
        breaks[0], breaks[-1] = lo - dx / 1000, hi + dx / 1000
    return np.searchsorted(breaks, x, side="left").astype(np.int64)      # (a, b] bins, 1-based


def indDiff_worker(x: dict) -> np.ndarray:
    """R/growthEnvironment.R:357-361: one chunk through the serial `diffChange` entry point."""
    return rcpp_api.diffChange(x["mat.in"], x["mat.out"], x["conc"], x["n_iter"])


def diffusion_niter(D, delta_time: float, field_size: float) -> np.ndarray:
    """Substeps per compound (R/growthEnvironment.R:380-382):
    ceiling( sqrt(D * 60 * 60 * deltaTime / (4*pi)) / fieldSize ), evaluated left to right."""
    D = np.asarray(D, dtype=np.float64)
    area = D * 60 * 60 * delta_time
    radius = np.sqrt(area / (4 * math.pi))
    return np.ceil(radius / field_size)


class GrowthEnvironment:
    """`concentrations` live on the GPU; `compound_D`, `conc_isConstant`, `fieldSize`, `fieldVol`,
    `mat_in`, `mat_out`, `field_pts` keep the reference's meaning (R/growthEnvironment.R:37-58)."""

    def __init__(self, mesh: FieldMesh, compounds, concentrations, compound_D, conc_isConstant=None,
                 device: int = 0, order: int = 0, exoenzymes=None, exoenzymes_conc=None):
        self.mesh = mesh
        self.field_pts = mesh.field_pts
        self.mat_in, self.mat_out = mesh.mat_in, mesh.mat_out
        self.fieldSize, self.fieldVol = mesh.field_size, mesh.field_vol
        self.fieldLayers, self.nfields = mesh.field_layers, mesh.nfields
        self.compounds = list(compounds)
        n_c = len(self.compounds)
        self.compound_D = np.asarray(compound_D, dtype=np.float64).reshape(n_c)
        self.conc_isConstant = (np.zeros(n_c, dtype=bool) if conc_isConstant is None
                                else np.asarray(conc_isConstant, dtype=bool).reshape(n_c))
        conc = np.asarray(concentrations, dtype=np.float64)
        if conc.shape != (self.nfields, n_c):
            raise ValueError("concentrations must be nfields x length(compounds)")
        self.plan = Plan(self.mat_in, self.mat_out, n_points=self.nfields, device=device, order=order)
        self.field = Field(self.plan, n_c)
        self.field.upload(conc)
        self.cells = Cells(self.plan)
        self.device = device
        # exoenzymes / exoenzymes.conc (R/growthEnvironment.R:55-56): nM, one column per enzyme
        self.exoenzymes = list(exoenzymes or [])
        self.exo_field = None
        if self.exoenzymes:
            ec = (np.zeros((self.nfields, len(self.exoenzymes))) if exoenzymes_conc is None
                  else np.asarray(exoenzymes_conc, dtype=np.float64))
            if ec.shape != (self.nfields, len(self.exoenzymes)):
                raise ValueError("exoenzymes_conc must be nfields x length(exoenzymes)")
            self.exo_field = Field(self.plan, len(self.exoenzymes))
            self.exo_field.upload(ec)

    # -- the `concentrations` slot, on demand ---------------------------------------------------
    @property
    def concentrations(self) -> np.ndarray:
        return self.field.download()

    @property
    def exoenzymes_conc(self) -> np.ndarray:
        return self.exo_field.download() if self.exo_field is not None else np.zeros((self.nfields, 0))

    def _diffuse_field(self, fld: Field, D: np.ndarray, is_constant, deltaTime: float):
        mn, mx, mean = fld.column_stats()
        with np.errstate(invalid="ignore"):
            variable = (mx > mn) & (D > 0)
        if is_constant is not None:
            variable &= ~is_constant
        ind_variable = np.flatnonzero(variable) + 1
        is_inf = np.isinf(D[ind_variable - 1])
        ind_inf = ind_variable[is_inf]
        ind_variable = ind_variable[~is_inf]
        if ind_variable.size > 0:
            niter = diffusion_niter(D[ind_variable - 1], deltaTime, self.fieldSize)
            fld.diffuse(niter, ind_variable)
        for c in ind_inf:
            fld.fill_column(int(c), float(mean[c - 1]))

    def exoenzyme_activity(self, deltaTime: float):
        """R/run_simulation.R:337-371: catalysis of every exoenzyme (closed-form Michaelis-Menten via
        lambertW0), then first-order decay; both on the device."""
        for k, ex in enumerate(self.exoenzymes, start=1):
            met_inds = [self.compounds.index(m) + 1 for m in ex.mets]   # match(exec@mets, compounds)
            self.field.exoenzyme_catalysis(self.exo_field, k, ex.Kcat, ex.Km, met_inds, ex.stoich,
                                           self.conc_isConstant, deltaTime)
        for k, ex in enumerate(self.exoenzymes, start=1):
            self.exo_field.scale_column(k, math.exp(-ex.lambda_ * deltaTime))
        return self

    def diffuse_compounds(self, deltaTime: float, cl=None, n_cores: int = 1):
        """The legacy dispatcher R/growthEnvironment.R:280-355: selection `sd > 0 & D > 0` (no
        isConstant / Inf rules), columns split into chunks with `cut`, every chunk through the
        serial `diffChange` entry point with the selected columns only (`indDiff_worker`)."""
        def run(fld, D):
            mn, mx, _ = fld.column_stats()
            with np.errstate(invalid="ignore"):
                ind_variable = np.flatnonzero((mx > mn) & (D > 0)) + 1
            return ind_variable
        ind_variable = run(self.field, self.compound_D)
        if ind_variable.size > 0:
            n_chunks = int(n_cores)
            if ind_variable.size / n_chunks < 3:
                n_chunks = int(math.ceil(n_chunks / 2))
            codes = chunk_codes(ind_variable.size, n_chunks) if n_chunks > 1 else np.ones(ind_variable.size, np.int64)
            for code in np.unique(codes):
                x = ind_variable[codes == code]
                chunk = {"mat.in": self.mat_in, "mat.out": self.mat_out, "conc": self.field.download(x),
                         "n_iter": diffusion_niter(self.compound_D[x - 1], deltaTime, self.fieldSize)}
                self.field.upload(indDiff_worker(chunk), cols=x)
        if self.exo_field is not None:
            exec_D = np.array([e.D for e in self.exoenzymes], dtype=np.float64)
            ind_variable = run(self.exo_field, exec_D)
            if ind_variable.size > 0:
                chunk = {"mat.in": self.mat_in, "mat.out": self.mat_out, "conc": self.exo_field.download(ind_variable),
                         "n_iter": diffusion_niter(exec_D[ind_variable - 1], deltaTime, self.fieldSize)}
                self.exo_field.upload(indDiff_worker(chunk), cols=ind_variable)
        return self

    def diffuse_compoundsPar(self, deltaTime: float, cl=None, n_cores: int = 1):
        """R/growthEnvironment.R:367-401 for the compound matrix (`cl` is unused there as well).

        Column selection `sd > 0 & D > 0 & !isConstant` (:372) is evaluated on the device as
        `max > min`; D = Inf columns are replaced by their mean (:373-374, :393-401); the others
        get `diffusion_niter` substeps through the diffChangePar contract (:385-390)."""
        self._diffuse_field(self.field, self.compound_D, self.conc_isConstant, deltaTime)
        # exoenzymes: same rule without the isConstant flag (R/growthEnvironment.R:406-434)
        if self.exo_field is not None:
            exec_D = np.array([e.D for e in self.exoenzymes], dtype=np.float64)
            self._diffuse_field(self.exo_field, exec_D, None, deltaTime)
        return self
