"""ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for `spl.fem.tensor.TensorFemSpace`
(used by sources/matrix_assembler.py:84-99: `.vector_space`, `.spaces`)."""
from spl.linalg.stencil import StencilVectorSpace


class TensorFemSpace:
    def __init__(self, *spaces, comm=None):
        self.spaces = list(spaces)
        npts = [S.nbasis for S in spaces]
        pads = [S.degree for S in spaces]
        self.vector_space = StencilVectorSpace(npts, pads, [False] * len(spaces))
