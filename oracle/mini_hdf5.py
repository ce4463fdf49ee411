"""TEST INFRASTRUCTURE ONLY -- the NetCDF-4 / HDF5 fixture reader now lives in the product package
(``metsim_b200/hdf5_min.py``: the minimal driver reads the same files); the fixture generator and the
oracle tests keep this name."""
from metsim_b200.hdf5_min import read_hdf5  # noqa: F401
