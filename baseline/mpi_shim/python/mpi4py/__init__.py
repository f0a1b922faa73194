"""A sliver of mpi4py over the shared-memory MPI shim (baseline/mpi_shim): just enough for the reference's
interoperability path (`mpi4torch.comm_from_mpi4py`, reference src/__init__.py:247-261, tests/test_mpi4pyinterop.py)
and for this library's `comm_from_mpi4py`.  Not a general mpi4py replacement."""
__version__ = "0.0-shim"
