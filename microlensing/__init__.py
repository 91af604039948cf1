"""Import-name shim: ``import microlensing`` resolves to the B200-native implementation, so
code written against ArnaudCassan/microlensing (``from microlensing import multipoles``,
``from microlensing import Rkp``; reference README.md:25-26,35-36) runs unchanged."""
import sys as _sys

import microlensing_b200 as _impl
from microlensing_b200 import version_info, __version__, multipoles, Rkp  # noqa: F401

_sys.modules[__name__ + ".multipoles"] = _impl.multipoles
_sys.modules[__name__ + ".multipoles.multipoles"] = _impl.multipoles.multipoles
_sys.modules[__name__ + ".multipoles.Rkp"] = _impl.multipoles.Rkp
_sys.modules[__name__ + ".Rkp"] = _impl.multipoles.Rkp
