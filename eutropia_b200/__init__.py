"""eutropia_b200 -- B200 (sm_100a) implementation of Eutropia's per-timestep compound diffusion and
cell<->field exchange path, behind the reference's own call surface.

  rcpp_api     diffChange / diffChangePar           (one-shot tier; R/RcppExports.R:4-10)
  session      Plan / Field / Cells                 (resident tier over the C ABI)
  environment  GrowthEnvironment.diffuse_compoundsPar (R/growthEnvironment.R:367-437)
  mesh         field points + mat.in / mat.out in the reference's format (setup, host)
  synth        synthetic inputs for tests and the bench

All compute is in libeutropia_b200.so (eutropia_b200/csrc); there is no CPU fallback.
"""
from . import _lib
from ._lib import EutropiaError, EutropiaIndexError
from .rcpp_api import diffChange, diffChangePar
from .session import Cells, Field, Plan
from .environment import GrowthEnvironment, diffusion_niter
from . import mesh, synth

__all__ = ["diffChange", "diffChangePar", "Plan", "Field", "Cells", "GrowthEnvironment",
           "diffusion_niter", "mesh", "synth", "EutropiaError", "EutropiaIndexError"]
__version__ = "0.1.0"
