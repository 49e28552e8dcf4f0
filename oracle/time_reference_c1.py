#!/usr/bin/env python
"""ORACLE / TEST INFRASTRUCTURE ONLY -- Tier-A CPU baseline: wall time of the UNMODIFIED reference
(/root/reference/sources/{kron_product,solvers,matrix_assembler}.py over the serial stand-ins in
oracle/shim/) on BASELINE config C1 (2-D, p = 3, 64 x 64 elements, n = 67 per axis, 4489 DOF):
  * kron_dot_v2 (sources/kron_product.py:56-89), the Kronecker mat-vec;
  * StencilMatrix.dot of the assembled operator (what every solver calls);
  * pcg + damped_jacobi, tol 1e-6, maxiter 10 (the pre-smoother of sources/mg_jac.py:87), b = 1;
  * the whole two-grid cycle of sources/mg_jac.py:85-119 (nc = 11: 8 x 8 coarse elements).
The reference is single-threaded Python per MPI rank; this runs one rank on one core.  It cannot
travel to the GPU box (no /root/reference there): run it in the dev container and commit the output
(profiles/r02_reference_tierA_c1.txt).    python oracle/time_reference_c1.py
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import numpy as np  # noqa: E402
import gen_golden as gg  # noqa: E402  (imports the reference over the shim; writes nothing on import)

p, nf, nc = 3, 67, 11
print("host: %d logical cores, 1 used (reference = serial Python); numpy %s" % (os.cpu_count(), np.__version__))
Tf = gg.make_open_knots(p, nf)
t0 = time.perf_counter()
S, A = gg.fem_problem(p, None, knots=list(Tf))
t_asm = time.perf_counter() - t0
V = S.vector_space
n1, n2 = V.npts
dof = n1 * n2
print("C1: p=%d n=%dx%d DOF=%d; reference assembly_2d: %.2f s" % (p, n1, n2, dof, t_asm))
# Kronecker mat-vec on the reference's own fixture matrices (utils.populate_1d_matrix), size of C1
A1, B1 = gg.mat1d(n1, p), gg.mat1d(n2, p)
gg.utils.populate_1d_matrix(A1, 5.0)
gg.utils.populate_1d_matrix(B1, 6.0)
X = gg.vec([n1, n2], [p, p], np.ones((n1, n2)))
reps = 3
t0 = time.perf_counter()
for _ in range(reps):
    Y = gg.kron_product.kron_dot_v2(A1, B1, X)
t = (time.perf_counter() - t0) / reps
print("kron_dot_v2            : %8.4f s per mat-vec = %.3e DOF/s" % (t, dof / t))
b = gg.StencilVector(V)
b[0:n1, 0:n2] = 1.0
t0 = time.perf_counter()
for _ in range(reps):
    y = A.dot(b)
t = (time.perf_counter() - t0) / reps
print("StencilMatrix.dot (shim): %8.4f s per mat-vec = %.3e DOF/s" % (t, dof / t))
t0 = time.perf_counter()
(x, info), dots = gg.logged(gg.solvers.pcg, A, gg.solvers.damped_jacobi, b, tol=1e-6, maxiter=10)
t = time.perf_counter() - t0
print("pcg + damped_jacobi (10 it.): %8.3f s, niter %d, res_norm %.3e => %.3e DOF/s per smoothing call"
      % (t, info["niter"], info["res_norm"], dof / t))
t0 = time.perf_counter()
out = gg.two_grid(p, nf, nc, "jac")
t = time.perf_counter() - t0
print("two-grid cycle mg_jac (incl. assembly_2d %.2f s): %8.3f s => %.3e DOF/s per cycle (without assembly: %.3e)"
      % (t_asm, t, dof / t, dof / max(t - t_asm, 1e-9)))
