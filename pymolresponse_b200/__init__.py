"""pymolresponse_b200: the linear-response (CPHF) hot path of berquist/pymolresponse, rebuilt for
NVIDIA B200 (sm_100a).

Same Python API as the reference for this path -- ``ao2mo.AO2MO``, ``explicit_equations_partial`` /
``explicit_equations_full``, ``operators.Operator``, ``solvers.ExactInv`` / ``ExactInvCholesky`` /
``IterativeLinEqSolver``, ``cphf.CPHF``, ``properties.electric.Polarizability``,
``properties.magnetic.Magnetizability`` -- with the arithmetic in hand-written CUDA behind the C ABI
of ``include/pmr_b200.h`` (``libpmr_b200.so``).  There is no CPU fallback.
"""

__version__ = "0.1.0"
