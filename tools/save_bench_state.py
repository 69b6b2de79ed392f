"""Writes tests/golden/bench_state_g<g>_k<outer>.npz on a B200: the APD state of the benchmarked SsN step (bench.py:
outer iteration 30, SsN step 1 of the 128x128-grid Class 1 solve) in a form the CPU arm can rebuild WITHOUT the
product library -- the nonzeros of the plans xk and vk (sparse at this state), the duals lk, the scalars ak, bk, bk1,
tk -- plus what the device step computes there (E, nnz(H0), components, W-cycles, accepted ll, |F| before / after),
which the oracle's step must reproduce (oracle/bench_step.py; bench.py --impl reference).

    python tools/save_bench_state.py [g=128] [outer=30]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ssnamg  # noqa: E402
drv = ssnamg.driver


def main():
    g = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    outer = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    m = n = g * g
    P = ssnamg.problems.grid_problem(g, seed=0)
    ssnamg.rng_reset()
    st = drv.capture_state(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], outer=outer, ssn_it=1, keep_plans=True)
    ssnamg.rng_reset()
    lk_new, Fk_new, info = drv.ssn_step(st)
    rec = {"g": g, "m": m, "n": n, "k": st["k"], "ssn_it": st["ssn_it"], "ak": st["ak"], "bk": st["bk"], "bk1": st["bk1"], "tk": st["tk"],
           "lk": st["lk"].cpu().numpy(), "xk_idx": st["xk_idx"], "xk_val": st["xk_val"], "vk_idx": st["vk_idx"], "vk_val": st["vk_val"],
           "expect_E": int(info["E"]), "expect_nnzH": int(info["nnzH"]), "expect_components": int(info["info"][0]),
           "expect_itamg": int(info["itamg"]), "expect_ll": int(info["ll"]), "expect_Fk_old_norm": info["Fk_old_norm"],
           "expect_Fk_new_norm": info["Fk_new_norm"], "expect_lk_new": lk_new.cpu().numpy(),
           "expect_wlk": st["wlk"].cpu().numpy()}
    # the fixture must reproduce wk: rebuild it the way oracle/bench_step.py does and compare on the device
    xk = torch.zeros(m * n, dtype=torch.float64, device="cuda"); xk[torch.from_numpy(rec["xk_idx"]).cuda()] = torch.from_numpy(rec["xk_val"]).cuda()
    wk = xk.clone(); wk[torch.from_numpy(rec["vk_idx"]).cuda()] += st["ak"] * torch.from_numpy(rec["vk_val"]).cuda()
    wk *= st["bk"] / st["ak"] ** 2; wk -= torch.from_numpy(P["c"]).cuda()
    err = float((wk - st["wk"]).abs().max() / st["wk"].abs().max())
    rec["wk_rebuild_maxrel"] = err
    out = os.path.join(ROOT, "tests", "golden", f"bench_state_g{g}_k{outer}.npz")
    np.savez_compressed(out, **rec)
    print("wrote", out, os.path.getsize(out), "bytes; nnz(xk)", rec["xk_idx"].size, "nnz(vk)", rec["vk_idx"].size, "E", rec["expect_E"],
          "ll", rec["expect_ll"], "cycles", rec["expect_itamg"], "wk rebuild max rel err", err)
    if os.path.isdir(os.path.join(ROOT, "gpurun_out")):
        import shutil
        shutil.copy(out, os.path.join(ROOT, "gpurun_out", os.path.basename(out)))


if __name__ == "__main__":
    main()
