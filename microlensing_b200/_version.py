"""Release number of the B200-native package.

Kept equal to the release of ArnaudCassan/microlensing whose ``multipoles`` path it replaces, so that
``microlensing.__version__`` reads the same for code that checks it."""
__version__ = "0.4.0"
version_info = tuple(int(part) for part in __version__.split("."))
