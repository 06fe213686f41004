"""Host-side mirror of the `growthEnvironment` slots the hot path reads and of its dispatcher
`diffuse_compoundsPar` (R/growthEnvironment.R:37-58, :367-437), with the field resident on a B200.

The dispatcher logic is R code in the reference and "stays in R" (SURVEY.md 8a-4); it is mirrored
here so that the whole round -- column selection, substep counts, D = Inf rule, diffusion -- can be
driven and tested without R, and so that a resident session never has to pull the field to the host
just to decide which columns diffuse.
"""
from __future__ import annotations

import math

import numpy as np

from .mesh import FieldMesh
from .session import Cells, Field, Plan

__all__ = ["GrowthEnvironment", "diffusion_niter"]


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
                 device: int = 0, order: int = 0):
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

    # -- the `concentrations` slot, on demand ---------------------------------------------------
    @property
    def concentrations(self) -> np.ndarray:
        return self.field.download()

    def diffuse_compoundsPar(self, deltaTime: float, cl=None, n_cores: int = 1):
        """R/growthEnvironment.R:367-401 for the compound matrix (`cl` is unused there as well).

        Column selection `sd > 0 & D > 0 & !isConstant` (:372) is evaluated on the device as
        `max > min`; D = Inf columns are replaced by their mean (:373-374, :393-401); the others
        get `diffusion_niter` substeps through the diffChangePar contract (:385-390)."""
        mn, mx, mean = self.field.column_stats()
        with np.errstate(invalid="ignore"):
            variable = (mx > mn) & (self.compound_D > 0) & (~self.conc_isConstant)
        ind_variable = np.flatnonzero(variable) + 1
        is_inf = np.isinf(self.compound_D[ind_variable - 1])
        ind_inf = ind_variable[is_inf]
        ind_variable = ind_variable[~is_inf]
        if ind_variable.size > 0:
            niter = diffusion_niter(self.compound_D[ind_variable - 1], deltaTime, self.fieldSize)
            self.field.diffuse(niter, ind_variable)
        for c in ind_inf:
            self.field.fill_column(int(c), float(mean[c - 1]))
        return self
