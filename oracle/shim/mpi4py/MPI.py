"""ORACLE / TEST INFRASTRUCTURE ONLY.  Serial (1-rank) stand-in for `mpi4py.MPI`.

Covers exactly what the reference touches: COMM_WORLD.Get_rank/Get_size/Barrier/allreduce,
Wtime, SUM, DOUBLE, and sub-communicator Allgatherv (reference call sites:
sources/kron_product.py:156,162,224,230; pyccel/pyccel_functions.py:158,164,228,235,242;
sources/mg_jac.py:95).
"""
import time
import numpy as np

SUM = "SUM"
DOUBLE = "DOUBLE"


def Wtime():
    return time.perf_counter()


class Comm:
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def Barrier(self):
        return None

    def allreduce(self, value, op=SUM):
        return value

    def Allgatherv(self, sendbuf, recvbuf):
        # recvbuf is either an array or [array, sizes, disps(, type)]; with one rank the
        # gathered line is the local segment.
        if isinstance(recvbuf, (list, tuple)):
            recvbuf = recvbuf[0]
        np.asarray(recvbuf)[...] = np.asarray(sendbuf).reshape(np.asarray(recvbuf).shape)

    def py2f(self):
        return 0


COMM_WORLD = Comm()
