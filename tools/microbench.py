"""Per-operator device timings (CUDA events, L2 flushed by plan-sized inputs) -- development aid."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    m = n = g * g
    peak = 6554.6
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(m * n, dtype=torch.float64, device="cuda", generator=gen)
    p = torch.ones(m, dtype=torch.float64, device="cuda"); q = torch.ones(n, dtype=torch.float64, device="cuda")
    lam = torch.randn(m + n, dtype=torch.float64, device="cuda", generator=gen) * 0.05
    gb = 8.0 * m * n / 1e9
    res = {}
    t, tmin = timeit(lambda: ssnamg.Ax(x, p, q)); res["Ax"] = (t, gb / t * 1e3)
    t, tmin = timeit(lambda: ssnamg.Aty(lam, p, q)); res["Aty"] = (t, gb / t * 1e3)
    w = x - 0.97
    t, _ = timeit(lambda: ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=("Axprox",))); res["prox_residual(Axprox)"] = (t, gb / t * 1e3)
    t, _ = timeit(lambda: ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=())); res["prox_residual(norm)"] = (t, gb / t * 1e3)
    t, _ = timeit(lambda: ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=("Axprox", "s"))); res["prox_residual(Axprox,s)"] = (t, (gb * 9 / 8) / t * 1e3)
    out = ssnamg.prox_residual(w, lam, p, q, 0.9, float("inf"), want=("s",))
    s = out["s"]; E = out["count"]
    t, _ = timeit(lambda: ssnamg.ASAt(s, p, q), reps=5); res[f"ASAt(E={E})"] = (t, (m * n * 2 / 1e9) / t * 1e3)
    y = torch.empty_like(x)
    t, _ = timeit(lambda: y.copy_(x)); res["torch copy (r+w)"] = (t, 2 * gb / t * 1e3)
    for k, (t, bw) in res.items():
        print(f"{k:32s} {t:9.3f} ms  {bw:9.1f} GB/s  {bw / peak:6.3f} of measured {peak}")
    # one inner solve on this synthetic active set
    H = ssnamg.ASAt(s, p, q)
    rhs = torch.randn(m + n, dtype=torch.float64, device="cuda", generator=gen)
    pd = {"bk1": 0.05, "tk": 0.9, "p": p, "q": q, "T": None, "H0": H, "z": rhs}
    opts = {"retol": 1e-11, "bigph": 1, "maxit": 30, "theta": 0.25, "smoth": 5, "cycle": "w", "isnsp": 1, "inter": 1, "guess": None}
    for rep in range(3):
        ssnamg.rng_reset(); l0 = ssnamg.launch_count(); torch.cuda.synchronize(); t0 = time.time()
        zeta, it, rres, info = ssnamg.Hybrid_AMG(pd, opts)
        torch.cuda.synchronize(); dt = time.time() - t0
        print(f"Hybrid_AMG: E={E} N={m + n} comps={info[0]} cycles={it} res={rres:.2e} {dt * 1e3:.2f} ms launches={ssnamg.launch_count() - l0}")


if __name__ == "__main__":
    main()
