"""Device timing of the screened line-search kernels (development aid / ncu target)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    rows = int(sys.argv[2]) if len(sys.argv) > 2 else g * g
    m, n = rows, g * g
    gen = torch.Generator(device="cuda").manual_seed(0)
    w = torch.rand(m * n, dtype=torch.float64, device="cuda", generator=gen) - 1.2
    p = torch.ones(m, dtype=torch.float64, device="cuda"); q = torch.ones(n, dtype=torch.float64, device="cuda")
    lam = torch.randn(m + n, dtype=torch.float64, device="cuda", generator=gen) * 0.05
    zeta = torch.randn(m + n, dtype=torch.float64, device="cuda", generator=gen) * 0.02
    for nt in (1, 32, 64):
        for _ in range(3):
            out = ssnamg.prox_trials_lin(w, lam, zeta, p, q, 0.9, 0.9, 1, nt)
        ssnamg.kernel_timer(True)
        for _ in range(10):
            ssnamg.prox_trials_lin(w, lam, zeta, p, q, 0.9, 0.9, 1, nt)
        ms, cnt = ssnamg.kernel_timer_read(); ssnamg.kernel_timer(False)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(10):
            ssnamg.prox_trials_lin(w, lam, zeta, p, q, 0.9, 0.9, 1, nt)
        torch.cuda.synchronize(); tot = (time.perf_counter() - t0) * 100
        print(f"m={m} n={n} nt={nt}: screen kernel {ms / cnt:.3f} ms ({8.0 * m * n / (ms / cnt) / 1e6:.0f} GB/s), whole batch {tot:.3f} ms, "
              f"candidates {float(out[nt]) / (m * n):.2e} of the entries")
    for _ in range(3):
        ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=())
    ssnamg.kernel_timer(True)
    for _ in range(10):
        ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=())
    ms, cnt = ssnamg.kernel_timer_read(); ssnamg.kernel_timer(False)
    print(f"single-trial kernel (norm only): {ms / cnt:.3f} ms")


if __name__ == "__main__":
    main()
