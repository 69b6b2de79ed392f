"""Device timing of the batched line-search kernel (plan_trials_kernel) -- development aid / ncu target."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    m = n = g * g
    gen = torch.Generator(device="cuda").manual_seed(0)
    w = torch.rand(m * n, dtype=torch.float64, device="cuda", generator=gen) - 0.97
    p = torch.ones(m, dtype=torch.float64, device="cuda"); q = torch.ones(n, dtype=torch.float64, device="cuda")
    lam = torch.randn(m + n, dtype=torch.float64, device="cuda", generator=gen) * 0.05
    for nt in (1, 2, 4, 8):
        lt = torch.stack([lam * (1 + 0.01 * t) for t in range(nt)]).contiguous()
        for _ in range(3):
            ssnamg.prox_trials(w, lt, p, q, 0.9)
        ssnamg.kernel_timer(True)
        for _ in range(10):
            ssnamg.prox_trials(w, lt, p, q, 0.9)
        ms, cnt = ssnamg.kernel_timer_read(); ssnamg.kernel_timer(False)
        t = ms / cnt
        print(f"trials NT={nt}: {t:.3f} ms/launch  {t / nt:.3f} ms/trial  {8.0 * m * n / t / 1e6:.0f} GB/s")
    for _ in range(3):
        ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=())
    ssnamg.kernel_timer(True)
    for _ in range(10):
        ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=())
    ms, cnt = ssnamg.kernel_timer_read(); ssnamg.kernel_timer(False)
    print(f"single-trial kernel (norm only): {ms / cnt:.3f} ms")


if __name__ == "__main__":
    main()
