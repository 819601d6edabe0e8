"""icqsol_b200 -- B200 (sm_100a) implementation of icqsol's boundary-element hot
path: dense Laplace single/double-layer matrix assembly over triangle pairs and
the dense solve that consumes it.  Same class and method names as the reference
(``icqsol.bem.icqLaplaceSolver.LaplaceSolver`` etc.); the arithmetic runs in
hand-written CUDA kernels behind a C ABI (include/icqb.h) loaded with ctypes.
There is no CPU path: without the compiled library and a CUDA device the
solvers raise.
"""
__version__ = '0.1.0'
