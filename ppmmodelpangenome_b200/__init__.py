"""ppmmodelpangenome_b200 -- B200-native likelihood engine for the PPM pangenome model.
Drop-in for the objective function of JasmineGamblin/PPMmodelPangenome (see include/ppm_b200.h)."""
from .engine import Inference, PPMError, measure_fp64_peak, lib, simulate_genes  # noqa: F401
