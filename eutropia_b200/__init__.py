"""eutropia_b200 -- B200 (sm_100a) implementation of Eutropia's per-timestep compound diffusion and
cell<->field exchange path, behind the reference's own call surface.

  rcpp_api     diffChange / diffChangePar           (one-shot tier; R/RcppExports.R:4-10)
  session      Plan / Field / Cells                 (resident tier over the C ABI)
  environment  GrowthEnvironment.diffuse_compoundsPar / diffuse_compounds / exoenzyme_activity
               (R/growthEnvironment.R:280-437, R/run_simulation.R:337-371)
  exchange     scavenge_compounds / adjust_uptake   (R/scavenge_compounds.R, R/adjust_uptake.R)
  recording    record_compounds / global_concentrations + RDS writer (R/run_simulation.R:480-526)
  mesh         field points + mat.in / mat.out in the reference's format (setup, host)
  synth        synthetic inputs for tests and the bench

All compute is in libeutropia_b200.so (eutropia_b200/csrc); there is no CPU fallback.
"""
from . import _lib
from ._lib import EutropiaError, EutropiaIndexError
from .rcpp_api import diffChange, diffChangePar
from .session import Cells, Field, Plan, SlabCells, SlabField
from .environment import Exoenzyme, GrowthEnvironment, diffusion_niter
from .exchange import adjust_uptake, scavenge_compounds
from . import mesh, recording, synth

__all__ = ["diffChange", "diffChangePar", "Plan", "Field", "Cells", "SlabField", "SlabCells", "GrowthEnvironment", "Exoenzyme",
           "diffusion_niter", "adjust_uptake", "scavenge_compounds", "mesh", "recording", "synth", "EutropiaError", "EutropiaIndexError"]
__version__ = "0.1.0"
