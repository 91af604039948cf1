version_info = (0, 4, 0)          # tracks the reference's microlensing/_version.py:1
__version__ = '.'.join(map(str, version_info))
