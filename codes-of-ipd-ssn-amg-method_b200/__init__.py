"""codes-of-ipd-ssn-amg-method_b200 -- B200-native (sm_100a CUDA) semismooth-Newton inner linear
solve of the IPD-SsN-AMG optimal-transport method.

The directory name is not a Python identifier; import it through the ``ssnamg`` shim at the
repository root (``import ssnamg``) or ``importlib.import_module``.  ``csrc/`` holds the CUDA
kernels and the C ABI (``include/ssnamg.h``); ``api.py`` mirrors the reference's MATLAB function
signatures on top of it; ``problems.py`` has the synthetic configurations of BASELINE.json.
"""
from . import problems, driver, matio                                            # noqa: F401
from ._lib import LIB_PATH, SIGNATURES, SsnError, load                    # noqa: F401
from .api import (APD_SsN_Class1, APD_SsN_Class2, warmup_class2, apd_begin_pot, apd_end_pot, ssn_step_class2, Ax, Aty, ASAt, ASAtz, ASAt_coo, active_coo, invAAt, invHHt, prox_residual, prox_residual_pot, prox_trials, prox_trials_lin, trial_vectors, linesearch, warmup_class1, warm_stage, apd_begin, apd_end,    # noqa: F401
                  strength, mis_set, cf_split, transfer, amg_setup, amg_clear,
                  MG_Vcycle, MG_Wcycle, Class_AMG, twogrid_bigph, twogrid, PCG, components, Hybrid_AMG, Hybrid_twogrid, ssn_step_class1,
                  aug_PCG, AMG4POT, PCG4POT, rescaled_system, jk_system, set_device_setup, set_fused_setup, set_cluster_solve, set_spgemm_slab_limit, spmv, spgemm,
                  transpose, DeviceCSR, rng_reset, rng_drawn, rand, launch_count, profile, set_dense_tail, set_persistent, kernel_timer, kernel_timer_read,
                  profile_dump)
