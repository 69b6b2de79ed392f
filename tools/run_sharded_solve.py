"""Row-sharded Class 1 solve on synthetic grid problems (BASELINE.json configs 4 and 5), one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
      tools/run_sharded_solve.py --grid 256 --max-outer 12

Every rank generates its own row slab of the cost on the device (the 256 x 256 cost is 34 GB), the
marginals are replicated.  Rank 0 prints one JSON line: warm-start / loop times, KKT and objective
histories, SsN / AMG / line-search counts, per-phase times and peak memory."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssnamg  # noqa: E402
from importlib import import_module  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=128)
    ap.add_argument("--max-outer", type=int, default=None)
    ap.add_argument("--max-seconds", type=float, default=None)
    ap.add_argument("--warm-maxit", type=int, default=100)
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--inner-solver", type=int, default=4, help="4 = Hybrid_AMG (reference default), 5 = Hybrid_twogrid")
    ap.add_argument("--stop-on-amg-divergence", action="store_true", help="end the solve when a W-cycle solve diverges (the reference carries on)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    sd = import_module("codes-of-ipd-ssn-amg-method_b200.sharded_driver")
    sharded = import_module("codes-of-ipd-ssn-amg-method_b200.sharded")
    g = args.grid; m = n = g * g
    r, l = ssnamg.problems.grid_marginals(g, seed=0, balanced=True)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()
    r0, r1 = sharded.row_range(rank, world, m)
    t0 = time.time()
    c_loc = sd.grid_cost_slab(g, r0, r1)
    torch.cuda.synchronize(); t_gen = time.time() - t0
    ssnamg.rng_reset()
    res = sd.APD_SsN_Class1_sharded(c_loc, dev(r), dev(l), dev(np.ones(m)), dev(np.ones(n)), rank, world,
                                    dist=dist if world > 1 else None, warm_maxit=args.warm_maxit, max_outer=args.max_outer,
                                    max_seconds=args.max_seconds, verbose=args.verbose, inner_solver=args.inner_solver,
                                    stop_on_amg_divergence=args.stop_on_amg_divergence)
    st = res["stats"]
    peak = torch.cuda.max_memory_allocated() / 2 ** 30
    # ---- slab kernels of the plan operators at this size (CUDA events around the launches, rank-local)
    opb = {}
    m_loc = r1 - r0
    p_loc = torch.ones(m_loc, dtype=torch.float64, device="cuda"); qd = torch.ones(n, dtype=torch.float64, device="cuda")
    lam = torch.randn(n + m_loc, dtype=torch.float64, device="cuda") * 0.01
    zeta = torch.randn(n + m_loc, dtype=torch.float64, device="cuda") * 0.01
    w = c_loc - 1.0
    slab_bytes = 8.0 * m_loc * n

    def ktime(fn, reps=5):
        for _ in range(2):
            fn()
        ssnamg.kernel_timer(True)
        for _ in range(reps):
            fn()
        ms, cnt = ssnamg.kernel_timer_read(); ssnamg.kernel_timer(False)
        if cnt == 0:
            raise RuntimeError("this operator's kernel is not bracketed by the kernel timer")
        return ms / cnt
    for name, fn in (("Ax", lambda: ssnamg.Ax(w, p_loc, qd)),
                     ("prox_residual(Axprox)", lambda: ssnamg.prox_residual(w, lam, p_loc, qd, 0.9, float("inf"), want=("Axprox",))),
                     ("trials_screen(64 steps)", lambda: ssnamg.prox_trials_lin(w, lam, zeta, p_loc, qd, 0.9, 0.9, 1, 64))):
        try:
            ms = ktime(fn)
            opb[name] = {"ms": ms, "GBps_per_gpu": slab_bytes / ms / 1e6, "GBps_all_gpus": world * slab_bytes / ms / 1e6}
        except Exception as e:                                          # keep the solve's record even if a microbenchmark fails
            opb[name] = {"error": str(e)[:200]}
    del w
    if rank == 0:
        out = {"config": f"grid{g}x{g}_vs_{g}x{g}_m{m}_n{n}", "n_gpus": world, "inner_solver": args.inner_solver, "plan_entries": m * n,
               "slab_rows": r1 - r0, "slab_GB_per_plan_vector": 8.0 * (r1 - r0) * n / 1e9, "cost_gen_s": t_gen,
               "warmup_s": res["warmup_seconds"], "loop_s": res["seconds"], "outer_its": res["outer_its"],
               "converged": bool(st["converged"]), "rel_kkt": res["rel_kkt"], "objective": res["fxk"][-1],
               "fxk": res["fxk"], "KKT_xk": res["KKT_xk"], "KKT_lk": res["KKT_lk"],
               "ssn_steps": int(sum(st["ssn_its"])), "ssn_its": st["ssn_its"], "amg_calls": st["amg_calls"],
               "amg_cycles": [int(v) for it in st["lin_its"] for v in it],
               "line_search_trials": st["ls_trials"], "line_search_passes": st["ls_passes"],
               "phase_ms": {"plan_wide_kernels_and_collectives": st["plan_ms"], "asat_assembly": st["asat_ms"], "hybrid_amg_replicated": st["solve_ms"]},
               "E_min_median_max": [int(np.min(st["E"])), int(np.median(st["E"])), int(np.max(st["E"]))] if st["E"] else None,
               "collectives": st["collectives"], "torch_peak_GB_rank0": peak, "steps": st["steps"], "amg_diverged": st["amg_diverged"], "slab_kernels_rank0": opb}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    main()
