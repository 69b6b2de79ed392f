"""Hybrid_AMG on a synthetic SsN system with the structure of an early-phase grid OT step: the active set is
s_ij = 1 iff |x_i - y_j| <= r grid cells (a disc neighbourhood), p = q = 1, bk1 = 0.5, tk = 2 (outer iteration 1).
Development aid: reproduces large (n+m = 131072) systems on ONE GPU without holding the 256 x 256 plan.
Usage: python tools/amg_synth.py g r [seed [log2 of the sparse-product slab limit]]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402


def main():
    g = int(sys.argv[1]); r = float(sys.argv[2]); seed = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    if len(sys.argv) > 4:
        ssnamg.set_spgemm_slab_limit(1 << int(sys.argv[4]))
    m = n = g * g
    idx = torch.arange(m, device="cuda")
    xa, xb = torch.div(idx, g, rounding_mode="floor").float(), (idx % g).float()
    s = torch.empty(m * n, dtype=torch.uint8, device="cuda")
    sv = s.view(n, m)                                        # column-major m x n: sv[j, i]
    step = max(1, (1 << 28) // m)
    for j0 in range(0, n, step):
        j1 = min(n, j0 + step)
        d2 = (xa[j0:j1, None] - xa[None, :]) ** 2 + (xb[j0:j1, None] - xb[None, :]) ** 2
        sv[j0:j1] = (d2 <= r * r).to(torch.uint8)
    E = int(s.sum(dtype=torch.int64))
    p = torch.ones(m, dtype=torch.float64, device="cuda"); q = torch.ones(n, dtype=torch.float64, device="cuda")
    H0 = ssnamg.ASAt(s, p, q)
    del s
    gen = torch.Generator(device="cuda").manual_seed(seed)
    z = torch.randn(m + n, dtype=torch.float64, device="cuda", generator=gen)
    pd = {"bk1": 0.5, "tk": 2.0, "q": q, "p": p, "T": None, "H0": H0, "z": z}
    opts = ssnamg.driver.CLASS1_AMG_OPTIONS
    ssnamg.profile(True)
    for name, solver in (("Hybrid_AMG", ssnamg.Hybrid_AMG), ("Hybrid_twogrid", ssnamg.Hybrid_twogrid)):
        for rep in range(2):
            ssnamg.rng_reset(); torch.cuda.synchronize(); t0 = time.perf_counter()
            zeta, it, res, info = solver(pd, opts)
            torch.cuda.synchronize()
            print(f"{name}: g={g} r={r} E={E} nnz(H0)={H0.nnz} comps={info[0]} its={it} res={res:.2e} |zeta|={float(torch.linalg.norm(zeta)):.3e} "
                  f"ms={(time.perf_counter() - t0) * 1e3:.1f}", flush=True)
    dump = ssnamg.profile_dump()
    print("\n".join(l for l in dump.splitlines() if "levels" in l or "solve." in l or "total" in l))


if __name__ == "__main__":
    main()
