"""``microlensing.multipoles`` surface: the five reference functions (GPU-backed) so that
``from microlensing import multipoles; multipoles.example()`` (reference README.md:25-26) works,
plus the batched entry points."""
from .multipoles import quadrupole, hexadecapole, _akl, _solvelenseq, example  # noqa: F401
from . import multipoles  # noqa: F401  (microlensing.multipoles.multipoles, the reference's real module path)
from . import Rkp  # noqa: F401
from ..batched import (magnification_points, magnification_map, magnification_models, models_chi2,  # noqa: F401
                       map_coordinates, lightcurve_coordinates, lightcurve_coordinates_at, alloc_outputs)
