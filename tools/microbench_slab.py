"""Kernel times of the plan-wide kernels on a row slab (rows x n), as a rank of an N-GPU run sees them: python tools/microbench_slab.py rows n"""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402

rows = int(sys.argv[1]); n = int(sys.argv[2])
gen = torch.Generator(device="cuda").manual_seed(0)
w = torch.rand(rows * n, dtype=torch.float64, device="cuda", generator=gen) - 0.97
p = torch.ones(rows, dtype=torch.float64, device="cuda"); q = torch.ones(n, dtype=torch.float64, device="cuda")
lam = torch.randn(n + rows, dtype=torch.float64, device="cuda", generator=gen) * 0.05
zeta = torch.randn(n + rows, dtype=torch.float64, device="cuda", generator=gen) * 0.01
flush = torch.empty(1 << 28, dtype=torch.uint8, device="cuda")
gb = 8.0 * rows * n / 1e9


def ktime(fn, reps=20):
    for _ in range(3):
        fn()
    ssnamg.kernel_timer(True)
    for _ in range(reps):
        flush.zero_()                                      # 256 MB: the slab (268 MB at 2048 x 16384) must not sit in L2
        fn()
    ms, cnt = ssnamg.kernel_timer_read(); ssnamg.kernel_timer(False)
    return ms / max(cnt, 1)


for name, fn in (("prox_residual(Axprox)", lambda: ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=("Axprox",))),
                 ("prox_residual(Axprox,s)", lambda: ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=("Axprox", "s"))),
                 ("trials_screen(64 steps)", lambda: ssnamg.prox_trials_lin(w, lam, zeta, p, q, 0.9, 0.9, 1, 64))):
    ms = ktime(fn)
    print(f"{name:28s} {ms * 1e3:8.1f} us  {gb / ms * 1e3:8.0f} GB/s  waves={os.environ.get('SSN_PLAN_WAVES', '2')} stage={os.environ.get('SSN_PLAN_STAGE', '1')}")
