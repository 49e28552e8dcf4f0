"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product (poms_b200/).

Single-process stand-in for the handful of `mpi4py.MPI` calls the reference's hot-path
modules make, so that `/root/reference/sources/*.py` import and run
unmodified in the dev container (no MPI is installed there).  One rank, no communication:
every collective degenerates to a copy.
"""
from . import MPI  # noqa: F401
