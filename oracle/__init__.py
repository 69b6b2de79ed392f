"""CPU oracle for the SsN inner linear solve -- TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy (+ a small C core, ``oracle/csrc/oracle_kernels.c``) restatement of the
reference's MATLAB functions on the hot path, one Python function per
reference ``.m`` function, same names, same argument meaning.  Every function cites the
reference ``file:line`` it follows (paths relative to the reference repository root).

PARITY UNPINNED: the reference ships no tests, golden vectors or expected outputs, and neither
MATLAB nor GNU Octave exists in the build container or on the GPU box, so this oracle cannot be
checked against the reference executing.  It is pinned instead by (i) the mathematical
identities the reference's own comments state (explicit ``A = [kron(I,p'); kron(q',I)]``,
``ASAt == A*diag(s)*A'``, adjointness, closed-form inverses), (ii) LP optimality of the full
solve on the reference's bundled input data, cross-checked with an independent LP solver, and
(iii) frozen conventions for everything that depends on MATLAB built-in internals (random
stream, sparse-product summation order, component ordering); see DESIGN.md.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product (``ssnamg``) never does.
"""
from .rng import MatlabRand, GLOBAL_STREAM, rand, rng_reset          # noqa: F401
from .plan_ops import Ax, Aty, ASAt, ASAtz, invAAt, invHHt, explicit_A  # noqa: F401
from .pcg import PCG                                                  # noqa: F401
from .amg import (strength, mis_set, cf_split, transfer, Class_AMG,   # noqa: F401
                  MG_Vcycle, MG_Wcycle, amg_state, twogrid_bigph, twogrid)
from .solvers import (components, Hybrid_AMG, Hybrid_twogrid, aug_PCG, AMG4POT,       # noqa: F401
                      PCG4POT)
