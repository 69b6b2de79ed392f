"""Cycles per grid-wide barrier: cooperative-groups grid.sync() vs the library's own barrier."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402
from importlib import import_module

lib = import_module("codes-of-ipd-ssn-amg-method_b200._lib")
ctx = lib.context()
for which, name in ((0, "cg grid.sync"), (1, "grid_barrier"), (0, "cg grid.sync"), (1, "grid_barrier"),
                    (2, "cluster barrier (release/acquire)"), (3, "global store + cluster barrier"), (4, "cluster barrier, relaxed arrive"),
                    (5, "store + barrier + ld.cg gather"), (6, "store + barrier + L1-cached gather"), (2, "cluster barrier (release/acquire)")):
    v = ctypes.c_double(0.0)
    ctx.call("ssn_debug_barrier_bench", 2000, which, ctypes.byref(v))
    print(f"{name:38s}: {v.value:8.0f} cycles per iteration", flush=True)

# building blocks of a pass of the DSMEM cluster solve kernel (amg_cluster.cu): 16 CTAs x 512 threads
for which, name in ((10, "z_sum1 (cluster-wide sum of one double)"), (18, "z_sum1_pull (own slot + 16 remote loads)"), (11, "z_barrier"), (16, "block reduction + barrier"), (17, "store + relaxed-arrive barrier"),
                    (412, "4 DSMEM gathers + barrier"), (812, "8 DSMEM gathers + barrier"), (1612, "16 DSMEM gathers + barrier"),
                    (819, "8 ld.shared::cluster gathers, OWN CTA"), (1619, "16 ld.shared::cluster gathers, OWN CTA"),
                    (820, "8 ld.shared::cluster gathers, next CTA"), (1620, "16 ld.shared::cluster gathers, next CTA"),
                    (821, "8 gathers, 3 of 4 own CTA"), (1621, "16 gathers, 3 of 4 own CTA"),
                    (822, "8 COALESCED remote loads, next CTA"), (1622, "16 COALESCED remote loads, next CTA"), (1623, "16 coalesced remote loads, 4 owners"),
                    (813, "8 DSMEM gathers, batches of 4"), (1613, "16 DSMEM gathers, batches of 4"),
                    (414, "4 L2 gathers + store + barrier"), (814, "8 L2 gathers + store + barrier"), (1614, "16 L2 gathers + store + barrier"),
                    (815, "8 local smem gathers + barrier"), (1615, "16 local smem gathers + barrier")):
    v = ctypes.c_double(0.0)
    ctx.call("ssn_debug_barrier_bench", 2000, which, ctypes.byref(v))
    print(f"{name:38s}: {v.value:8.0f} cycles per iteration", flush=True)
