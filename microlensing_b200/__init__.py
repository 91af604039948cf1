"""B200-native (sm_100a, FP64 CUDA) drop-in for the ``multipoles`` hot path of
ArnaudCassan/microlensing.  Importing the package does not touch the GPU; the first call
into the C ABI loads ``libmlmultipoles.so`` and creates a context (no CPU fallback)."""
from ._version import version_info, __version__  # noqa: F401
from . import multipoles  # noqa: F401
from .multipoles import Rkp  # noqa: F401
from .batched import (magnification_points, magnification_map, magnification_models, models_chi2,  # noqa: F401
                      map_coordinates, lightcurve_coordinates, lightcurve_coordinates_at, alloc_outputs)
