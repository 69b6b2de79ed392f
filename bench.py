#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200: one semismooth-Newton (SsN) inner solve
step of the SsN-AMG optimal-transport solver on the 128x128-vs-128x128 grid (m = n = 16384,
268M-entry fp64 plan), at a realistic APD state captured from the device-resident solve.

A "step" is one pass of the hot path (Class1/APD_SsN_Class1.m:137-212 of the reference):
  fused residual + active set (one read of the plan-sized wk)  ->  ASAt assembly  ->
  Hybrid_AMG (AMG setup + W-cycles to 1e-11)  ->  Armijo line-search trial(s)  ->  new residual.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torchrun)
  python bench.py --impl reference ...                     (the CPU oracle port on the host cores: the same step at full
                                                            size from tests/golden/bench_state_g128_k30.npz, one step, no GPU)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ssn_amg_inner_solve_step_time_128x128_grid"
UNIT = "ms/step"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, default=128, help="grid side g (m = n = g*g); 128 is the headline config")
    ap.add_argument("--state-outer", type=int, default=30, help="APD outer iteration whose first SsN step is benchmarked")
    ap.add_argument("--config", default="class1_128", choices=["class1_128", "class2_64"],
                    help="class1_128: the headline (BASELINE configs[3]); class2_64: one SsN step of partial OT on 64x64 grids (configs[2])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-solve", action="store_true", help="skip the timing of the whole Class1 solve (N=1 only)")
    return ap.parse_args()


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------ CPU oracle arm

FIXTURE = os.path.join(ROOT, "tests", "golden", "bench_state_g{g}_k{k}.npz")


def load_problems_module():
    """problems.py (NumPy only) loaded by path: the CPU arm never imports the product package."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_ssn_problems", os.path.join(ROOT, "codes-of-ipd-ssn-amg-method_b200", "problems.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    return mod


def cpu_step_full(host_state):
    """The oracle (CPU port of the reference; MATLAB/Octave are absent) on the SAME step at FULL size: the whole
    m x n plan, every line-search trial, nothing sampled or scaled (oracle/bench_step.py).  Returns
    (ms, info, sample text)."""
    from oracle import bench_step
    ms, lk_new, Fk_new, info = bench_step.timed_step(host_state)
    ph = info["phases_s"]
    txt = (f"oracle (NumPy/SciPy port of the reference; MATLAB/Octave absent), the WHOLE step once, nothing sampled or extrapolated: "
           f"{host_state['m']}x{host_state['n']} plan, residual + active set {ph['residual_s']:.1f} s, ASAt {ph['asat_s']:.1f} s, "
           f"Hybrid_AMG on the {host_state['m'] + host_state['n']}-node system {ph['hybrid_amg_s']:.1f} s (single-threaded SciPy), all "
           f"{info['ll'] + 1} Armijo trials {ph['line_search_s']:.1f} s, new residual {ph['new_residual_s']:.1f} s; plan-wide expressions "
           f"on {info['threads']} column blocks at the same time ({info['threads']} host threads)")
    return ms, lk_new, Fk_new, info, txt


def device_state_from_fixture(torch, ssnamg, path, P):
    """The benchmarked APD state on the device, from the fixture's sparse plans (Class1/APD_SsN_Class1.m:113-126)."""
    d = np.load(path)
    m, n = int(d["m"]), int(d["n"])
    ak, bk = float(d["ak"]), float(d["bk"])
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    xk = torch.zeros(m * n, dtype=torch.float64, device="cuda"); xk[dev(d["xk_idx"])] = dev(d["xk_val"])
    vk = torch.zeros(m * n, dtype=torch.float64, device="cuda"); vk[dev(d["vk_idx"])] = dev(d["vk_val"])
    p = dev(P["p"]); q = dev(P["q"]); c = dev(P["c"])
    wk, axk = ssnamg.apd_begin(c, xk, vk, p, q, ak, bk)                 # :125 and Ax(xk) of :126
    del xk, vk, c
    b = dev(np.concatenate([P["r"], P["l"]]))
    lk = dev(d["lk"])
    bk1, tk = float(d["bk1"]), float(d["tk"])
    wlk = bk1 * (lk - 1 / bk * (axk - b)) - b                           # :126
    return {"wk": wk, "lk": lk, "wlk": wlk, "bk1": bk1, "tk": tk, "k": int(d["k"]), "ssn_it": int(d["ssn_it"]), "p": p, "q": q,
            "gama": float("inf"), "E": int(d["expect_E"])}


def step_config(workload, info, n_ll):
    """What both arms report under `config`: the workload and the discrete facts of the step, which the device
    path and the CPU oracle compute independently (equal dictionaries = the two arms did the same work)."""
    return {"workload": workload, "E_active": int(info["E"]), "nnz_H0": int(info["nnzH"]), "amg_cycles": int(info["itamg"]),
            "components": int(info["components"]), "line_search_trials": int(n_ll) + 1,
            "l2_flush": "inputs larger than L2 (2.1 GB plan vector per pass)"}


# ------------------------------------------------------------------------------ main

def main():
    args = parse()
    # Only the final JSON line may reach stdout: libraries (NCCL's version banner, torch warnings) write
    # to file descriptor 1 directly, so it is pointed at stderr until the result is printed.
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return 0 if rank != 0 else (run_reference_class2(args) if args.config == "class2_64" else run_reference(args))
    if args.config == "class2_64":
        return 0 if rank != 0 else run_class2(args)
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1 and args.impl == "ours":
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import ssnamg
    drv = ssnamg.driver
    g = args.grid
    m = n = g * g
    peak, peak_src = peaks()

    # ---- the APD state of the benchmarked step: rebuilt on the device from the committed fixture (the same file the CPU
    # arm reads; tests/test_gpu_traces.py checks that the device solve passes through exactly this state); without the
    # fixture the solve is run up to that step.  Identical on every rank.
    t0 = time.time()
    P = ssnamg.problems.grid_problem(g, seed=0)
    fixture = FIXTURE.format(g=g, k=args.state_outer)
    if os.path.exists(fixture):
        state = device_state_from_fixture(torch, ssnamg, fixture, P)
        state_src = os.path.relpath(fixture, ROOT)
    else:
        ssnamg.rng_reset()
        state = drv.capture_state(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], outer=args.state_outer, ssn_it=1)
        state_src = "device solve run up to the step (fixture missing)"
    torch.cuda.synchronize()
    t_state = time.time() - t0
    full = None; full5 = None
    if world == 1 and not args.no_full_solve:
        # the whole Class 1 solve (the metric's "solve time") through the library's one-call entry point ssn_apd_ssn_class1
        ssnamg.rng_reset()
        full = ssnamg.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"])
        full.pop("xk", None); torch.cuda.empty_cache()
        # the same script with its other multilevel option, inner_solver = 5 (Hybrid_twogrid, Class1/APD_SsN_Class1.m:70,178)
        ssnamg.rng_reset()
        full5 = ssnamg.APD_SsN_Class1(P["c"], P["r"], P["l"], P["p"], P["q"], P["gama"], inner_solver=5)
        full5.pop("xk", None); torch.cuda.empty_cache()
    del P
    workload = f"grid{g}x{g}_vs_{g}x{g}_m{m}_n{n}_outer{state['k']}_ssn{state['ssn_it']}"

    if world > 1:
        from importlib import import_module
        shard = import_module("codes-of-ipd-ssn-amg-method_b200.sharded")
        step_fn = shard.make_sharded_step(state, rank, world)
    else:
        def step_fn():
            # ONE library call per step (ssn_ssn_step_class1): residual -> ASAt -> Hybrid_AMG -> line search -> residual
            ssnamg.rng_reset()
            return ssnamg.ssn_step_class1(state["wk"], state["lk"], state["wlk"], state["p"], state["q"], state["bk1"], state["tk"])

    if world > 1:
        k3_w, k3_lam, k3_p, k3_rows = step_fn.w_loc, step_fn._lam_loc(state["lk"]), step_fn.p_loc, step_fn.m_loc
        state["wk"] = None                                  # the full plan vector is not needed any more
        torch.cuda.empty_cache()
    else:
        k3_w, k3_lam, k3_p, k3_rows = state["wk"], state["lk"], state["p"], m

    lam8 = torch.stack([k3_lam + 0.9 ** t * 1e-3 for t in range(8)]).contiguous()      # eight trial vectors (rows)

    def kernel_ms(fn, reps=10):
        """Mean duration of the plan-wide kernel inside fn(), CUDA events on the launching stream
        around the launch itself (ssn_kernel_timer)."""
        for _ in range(3):
            fn()
        ssnamg.kernel_timer(True)
        for _ in range(reps):
            fn()
        ms, cnt = ssnamg.kernel_timer_read()
        ssnamg.kernel_timer(False)
        if cnt > 0 and ms > 0:
            return ms / cnt
        # the library's timer did not fire: CUDA events around the whole call (the plan-wide kernel + its small finish
        # kernels and the host read of the call, i.e. slightly pessimistic)
        print(f"bench: ssn_kernel_timer returned ({ms}, {cnt}); timing the calls with torch events instead", file=sys.stderr)
        a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a0.record()
        for _ in range(reps):
            fn()
        a1.record(); torch.cuda.synchronize()
        return a0.elapsed_time(a1) / reps

    for _ in range(max(args.warmup, 3)):
        lk_new, Fk_new, info = step_fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    l0 = ssnamg.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    prof_range = os.environ.get("SSN_BENCH_PROFILE") == "1"      # ncu --profile-from-start off: timed region only
    if prof_range:
        torch.cuda.profiler.start()
    e0.record()
    for _ in range(args.steps):
        lk_new, Fk_new, info = step_fn()
    e1.record()
    torch.cuda.synchronize()
    if prof_range:
        torch.cuda.profiler.stop()
    launches = ssnamg.launch_count() - l0
    if world > 1:
        dist.barrier()
        # one more step with device-synchronised phase laps (outside the timed region): plan-wide part / ASAt / AMG, max over ranks
        step_fn.profile = True
        _, _, info_p = step_fn()
        step_fn.profile = False
        tph = torch.tensor([info_p["ms_plan"], info_p["ms_asat"], info_p["ms_amg"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(tph, op=dist.ReduceOp.MAX)
        info = dict(info, ms_plan=float(tph[0]), ms_asat=float(tph[1]), ms_amg=float(tph[2]))
    ms_total = e0.elapsed_time(e1)
    if world == 1:
        # the same step through the operator-level calls (driver.ssn_step), outside the timed region: its phase laps
        # (each closed by a device synchronise) and the Newton direction the kernel timings below need
        for _ in range(3):                                   # the first calls of this path pay its one-time costs
            ssnamg.rng_reset()
            lk_py, _, info_py = drv.ssn_step(state)
        assert float((lk_py - lk_new).abs().max()) <= 1e-12 * float(lk_new.abs().max()), "one-call step and operator-level step disagree"
        info = dict(info_py, **{k: info[k] for k in ("E", "nnzH", "itamg", "ll", "ls_passes")})
    # dominant HBM-bound kernel, timed alone with CUDA events on the launching stream (sampler still running)
    k3_ms = kernel_ms(lambda: ssnamg.prox_residual(k3_w, k3_lam, k3_p, state["q"], state["tk"], float("inf"), want=("Axprox",)))
    tr_ms = kernel_ms(lambda: ssnamg.prox_trials(k3_w, lam8, k3_p, state["q"], state["tk"], float("inf")))
    # the kernel the line search of this step actually runs: the screened one, on the step's own direction
    zeta = info["zeta"]
    zeta_loc = step_fn._lam_loc(zeta) if world > 1 else zeta
    slots = max(1.0, float(k3_rows) * n)
    dens = float(ssnamg.prox_trials_lin(k3_w, k3_lam, zeta_loc, k3_p, state["q"], state["tk"], 0.9, 0, 1)[1]) / slots
    nt_lin = 64 if dens <= 0.10 else 16
    lin_call = lambda: ssnamg.prox_trials_lin(k3_w, k3_lam, zeta_loc, k3_p, state["q"], state["tk"], 0.9, 1, nt_lin)
    lin_ms = kernel_ms(lin_call)                            # plan_trials_screen_kernel alone (the plan-wide kernel of a batch)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        lin_call()                                          # synchronous: screen + scan + host read + compact + eval + finish
    torch.cuda.synchronize(); lin_batch_ms = (time.perf_counter() - t0) * 1e3 / 10
    sampler.stop_flag = True; sampler.join(timeout=2)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t[0])
    ms_step = ms_total / args.steps

    out = {"metric": METRIC, "value": ms_step, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "impl": "ours",
           "config": step_config(workload, dict(info, components=info["info"][0]), info["ll"]),
           "run": {"line_search_passes": int(info.get("ls_passes", 0)), "state_build_s": round(t_state, 1), "state_source": state_src,
                   "sharding": "plan rows over ranks; AMG replicated" if world > 1 else "single GPU"},
           "breakdown_ms": {"plan_wide_kernels_and_collectives": info.get("ms_plan"), "asat_assembly": info.get("ms_asat"),
                            "hybrid_amg_replicated": info.get("ms_amg"),
                            "note": "host-timed phases of the last step (each closed by a device synchronise): the plan-wide "
                                    "part is what row-sharding divides, the AMG solve is replicated on every rank"},
           "ssn_steps_per_s": 1e3 / ms_step, "gpu_launches": int(launches), "clocks": sampler.summary()}

    # the part of the step that row-sharding divides (north star: "near-linear 1->8 scaling on the sharded plan operators"):
    # the two fused residuals and the line-search reads of wk with their collectives -- its time, and the bytes of the
    # WHOLE plan those reads cover per second, over all ranks
    plan_reads = int(info.get("ls_passes", 0)) + 2
    if info.get("ms_plan"):
        out["plan_operators"] = {"ms_per_step": info["ms_plan"], "reads_of_the_plan_per_step": plan_reads,
                                 "aggregate_GBps": plan_reads * 8.0 * m * n / (info["ms_plan"] * 1e-3) / 1e9,
                                 "frac_of_n_gpus_x_peak": plan_reads * 8.0 * m * n / (info["ms_plan"] * 1e-3) / 1e9 / (peak * world),
                                 "note": "device-synchronised lap of one step (max over ranks), kernels + collectives + the host reads of the "
                                         "line search; the replicated AMG solve is not in it"}
    if "plan_operators" in out:
        # the same reads counted by their kernels alone (CUDA events around each launch, this rank's slab): what scales with
        # the number of GPUs; the lap above adds the collectives, the Python between the kernels and the host reads
        kern_ms = 2 * k3_ms + max(plan_reads - 2, 0) * lin_ms
        out["plan_operators"].update({"kernels_ms_per_step": kern_ms,
                                      "kernels_aggregate_GBps": plan_reads * 8.0 * m * n / (kern_ms * 1e-3) / 1e9,
                                      "kernels_frac_of_n_gpus_x_peak": plan_reads * 8.0 * m * n / (kern_ms * 1e-3) / 1e9 / (peak * world)})
    bytes_pass = 8.0 * k3_rows * n                          # one read of the (slab of the) plan-sized wk
    passes = int(info.get("ls_passes", 0))
    slab_txt = f", rank 0's {k3_rows}-row slab" if world > 1 else ""

    def roof(kernel, ms, launches):
        ach = bytes_pass / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "kernel": kernel + slab_txt, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_pass, "avg_launch_ms": ms,
                "launches_per_step": launches, "share_of_step": launches * ms / ms_step,
                "traffic": None, "frac_of_8TBps_nominal": ach / 8000.0}
    r_k3 = roof("plan_reduce_kernel<PROX> (fused SsN residual: z, prox, Ax(prox), ||prox||^2; one read of wk)", k3_ms, 2)
    r_tr = roof("plan_trials_kernel<NT=8> (dense fallback of the line search: 8 Armijo trials per read of wk, every entry evaluated)", tr_ms, 0)
    r_tr["note"] = ("not launched in this step; used when more than 25 % of the entries survive the screen.  Reads wk once for "
                    "8 trials, so its limiter is the fp64 pipe, not HBM (profiles/trials_full_r1.csv)")
    screened = dens <= 0.25
    r_lin = roof("plan_trials_screen_kernel (screened line search: one read of wk per batch of Armijo steps marks the entries where some "
                 "step of the batch can be active; ~7 fp64 operations per entry whatever the batch size)", lin_ms, max(passes - 1, 0) if screened else 0)
    r_lin["surviving_entries"] = dens
    r_lin["batch_steps"] = nt_lin
    r_lin["batch_ms_host_timed"] = lin_batch_ms
    r_lin["share_of_step"] = r_lin["launches_per_step"] * lin_batch_ms / ms_step
    r_lin["note"] = (f"a batch of {nt_lin} steps = this kernel + the candidate scan / compact / eval kernels on the surviving entries: "
                     f"{lin_batch_ms:.3f} ms in all (host-timed, synchronous call through the Python binding); share_of_step uses that figure")
    if world == 1:
        r_k3["traffic"] = load_traffic("k3"); r_tr["traffic"] = load_traffic("trials"); r_lin["traffic"] = load_traffic("screen")
    cands = sorted([r_lin, r_k3, r_tr], key=lambda r: -r["share_of_step"])
    dominant, other = cands[0], cands[1:]
    out["roofline"] = dominant
    if world == 1:
        out["dominant_by_time"] = amg_kernel_share(ssnamg, drv, state, ms_step)
    out["roofline_other"] = other
    if world > 1:
        out["collectives_per_step"] = int(info.get("collectives", 0)) // max(1, args.steps + max(args.warmup, 3))
    if world > 1:
        # ---- e2e on N GPUs: every rank copies ITS slab of wk (and the duals) from pinned host memory each step, runs the sharded
        # step on the copies and reads lk_new / Fk_new back; wall clock between barriers, max over ranks
        h_w = step_fn.w_loc.cpu().pin_memory(); h_lk = step_fn.lk.cpu().pin_memory(); h_wlk = step_fn.wlk.cpu().pin_memory()
        keep = (None, step_fn.lk, step_fn.wlk)
        step_fn.w_loc = None; k3_w = None; lin_call = None; torch.cuda.empty_cache()

        def host_step_sharded():
            step_fn.w_loc = h_w.cuda(non_blocking=True); step_fn.lk = h_lk.cuda(non_blocking=True); step_fn.wlk = h_wlk.cuda(non_blocking=True)
            a_, b_, _ = step_fn()
            return a_.cpu(), b_.cpu()
        for _ in range(2):
            host_step_sharded()
        torch.cuda.synchronize(); dist.barrier()
        ke = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ke):
            lk_h, Fk_h = host_step_sharded()
        torch.cuda.synchronize()
        te = torch.tensor([(time.perf_counter() - t0) * 1e3 / ke], dtype=torch.float64, device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        hb = torch.tensor([float(h_w.numel() * 8 + h_lk.numel() * 8 + h_wlk.numel() * 8)], dtype=torch.float64, device="cuda")
        dist.all_reduce(hb)
        out["e2e"] = {"value": float(te[0]), "unit": UNIT, "h2d_bytes_per_step": int(hb[0]),
                      "d2h_bytes_per_step": int(world * (lk_h.numel() * 8 + Fk_h.numel() * 8)),
                      "note": "bytes summed over the ranks: each rank uploads its own row slab over its own PCIe link"}
        step_fn.w_loc, step_fn.lk, step_fn.wlk = None, keep[1], keep[2]
    if world == 1:
        # ---- e2e: the same step through host buffers (pinned), H2D of the step's inputs + D2H of its result
        hstate = {k: (v.cpu().pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in state.items()}
        h2d = sum(hstate[k].numel() * hstate[k].element_size() for k in ("wk", "lk", "wlk", "p", "q"))

        def host_step():
            # the plugin call a caller with HOST arrays makes: ssn_ssn_step_class1_host copies wk, lk, wlk, p, q in and lk_new, Fk_new out
            ssnamg.rng_reset()
            return ssnamg.ssn_step_class1(hstate["wk"], hstate["lk"], hstate["wlk"], hstate["p"], hstate["q"], hstate["bk1"], hstate["tk"],
                                          host_call=True)
        for _ in range(2):
            host_step()
        torch.cuda.synchronize()
        ke = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(ke):
            lk_h, Fk_h, _ = host_step()
        torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1e3 / ke
        out["e2e"] = {"value": e2e_ms, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                      "d2h_bytes_per_step": int(lk_h.numel() * 8 + Fk_h.numel() * 8)}
        del hstate
        # ---- secondary figures named by the metric: PCG iterations/s on the full KKT system, W-cycles/s
        out.update(secondary_metrics(ssnamg, drv, state, m, n))
        if full is not None:
            fs = full["stats"]
            out["full_solve"] = {"total_s": full["seconds"] + full["warmup_seconds"], "loop_s": full["seconds"],
                                 "warmup_s": full["warmup_seconds"], "outer_its": full["outer_its"], "converged": bool(fs["converged"]),
                                 "rel_kkt": full["rel_kkt"], "objective": full["fxk"][-1], "ssn_steps": int(sum(fs["ssn_its"])),
                                 "entry_point": "ssn_apd_ssn_class1 (one library call: warm start + outer loop + SsN steps)",
                                 "line_search_trials": int(fs["ls_trials"]), "amg_solves": int(fs["amg_calls"]),
                                 "amg_s": fs["solve_s"], "plan_s": fs["plan_s"], "asat_s": fs["asat_s"],
                                 "status": ("converged to rel-KKT <= 1e-6" if fs["converged"] else
                                            f"STOPPED AT maxit = {full['outer_its']} outer iterations (Class1/APD_SsN_Class1.m:35), NOT converged: "
                                            f"rel-KKT {full['rel_kkt']:.1e} > KKT_Tol 1e-6"),
                                 "note": "Class1/APD_SsN_Class1.m with its own limits (maxit = 100 outer iterations, KKT_Tol 1e-6) and its default "
                                         "inner_solver = 4 (Hybrid_AMG); a time to the iteration cap is not a time to solution"}
        if full5 is not None:
            f5 = full5["stats"]
            out["full_solve_inner_solver5"] = {
                "total_s": full5["seconds"] + full5["warmup_seconds"], "loop_s": full5["seconds"], "warmup_s": full5["warmup_seconds"],
                "outer_its": full5["outer_its"], "converged": bool(f5["converged"]), "rel_kkt": full5["rel_kkt"], "objective": full5["fxk"][-1],
                "ssn_steps": int(sum(f5["ssn_its"])), "line_search_trials": int(f5["ls_trials"]), "inner_solves": int(f5["amg_calls"]),
                "inner_solve_s": f5["solve_s"], "plan_s": f5["plan_s"], "asat_s": f5["asat_s"],
                "note": "the same script with inner_solver = 5 (Hybrid_twogrid, the reference's two-level option, "
                        "Class1/APD_SsN_Class1.m:70,178): a time to solution when `converged` is true"}
        if not args.no_cpu_baseline:
            host_state = {"wk": state["wk"].cpu().numpy(), "lk": state["lk"].cpu().numpy(), "wlk": state["wlk"].cpu().numpy(),
                          "p": np.ones(m), "q": np.ones(n), "tk": state["tk"], "bk1": state["bk1"], "m": m, "n": n}
            del state["wk"]; torch.cuda.empty_cache()
            cpu_ms, lk_cpu, Fk_cpu, cinfo, cpu_txt = cpu_step_full(host_state)
            out["cpu_baseline"] = {"value": cpu_ms, "unit": UNIT, "cores": cinfo["threads"], "kind": "port", "sample": cpu_txt,
                                   "same_step_as_device": {"config_equal": step_config(workload, cinfo, cinfo["ll"]) == out["config"],
                                                           "lk_new_max_rel_diff": float(np.max(np.abs(lk_cpu - lk_new.cpu().numpy())) / np.max(np.abs(lk_cpu))),
                                                           "Fk_new_norm_cpu": cinfo["Fk_new_norm"], "Fk_new_norm_device": info["Fk_new_norm"]}}
    if rank == 0:
        emit(out)
    if world > 1:
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def emit(obj):
    """Prints the one JSON line on the real stdout."""
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(obj), flush=True)


def load_traffic(which):
    """dram bytes per launch of a plan-wide kernel from the committed ncu capture (profiles/), if any."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return float(d[which]["dram_bytes_per_launch"])
    except Exception:
        return None


def amg_kernel_share(ssnamg, drv, state, ms_step):
    """The kernel with the largest share of the step when the line search is short: Class_AMG's solve loop as ONE
    kernel.  It is latency-bound (about 115 dependent passes per W-cycle, each a few thousand cycles of gathers and
    one barrier), so it is reported by time, not by an HBM fraction.  Times from the library's phase
    profiler (host-timed, synchronised phases)."""
    import torch
    ev = ssnamg.prox_residual(state["wk"], state["lk"], state["p"], state["q"], state["tk"], float("inf"), want=("Axprox", "s"))
    H0 = ssnamg.ASAt(ev["s"], state["p"], state["q"])
    Fk = state["bk1"] * state["lk"] - ev["Axprox"] - state["wlk"]
    pd = {"bk1": state["bk1"], "tk": state["tk"], "q": state["q"], "p": state["p"], "T": None, "H0": H0, "z": -Fk}
    ssnamg.profile(True)
    reps = 3
    for _ in range(reps):
        ssnamg.rng_reset(); _, itamg, _, _ = ssnamg.Hybrid_AMG(pd, drv.CLASS1_AMG_OPTIONS)
    torch.cuda.synchronize()
    dump = ssnamg.profile_dump()
    ssnamg.profile(False)
    phases = {}
    for line in dump.splitlines():
        parts = line.split()
        if len(parts) >= 4 and parts[-2] == "calls" and "#launches" not in line:
            try:
                phases[" ".join(parts[:-3])] = float(parts[-3]) / reps
            except ValueError:
                pass
    names = (("solve.dsm_solve_kernel", "dsm_solve_kernel (Class_AMG.m:89-107 + MG_Wcycle.m inside one 16-CTA cluster, level vectors in distributed shared memory)",
              "latency (cluster barriers + ld.shared::cluster gathers)"),
             ("solve.cluster_solve_kernel", "cluster_solve_kernel (the same loop inside one cluster, vectors in global memory)", "latency (cluster barriers + L2 gathers)"),
             ("solve.persist_solve_kernel", "persist_solve_kernel (the same loop as one cooperative grid-wide kernel)", "latency (grid barriers)"))
    k, kname, bound = None, None, None
    for key, text, b in names:
        if phases.get(key):
            k, kname, bound = phases[key], text, b
            break
    return {"kernel": kname, "bound": bound,
            "ms_per_step": k, "share_of_step": (k / ms_step) if k else None, "wcycles": int(itamg),
            "ms_per_wcycle": (k / itamg) if k and itamg else None,
            "amg_setup_ms": phases.get("amg_setup total"), "dense_tail_build_ms": phases.get("solve.build_dense_tail"),
            "amg_solve_loop_ms": phases.get("class_amg solve loop total"),
            "note": "host-timed phases of the library's profiler, each closed by a device synchronise (so slightly above their share of the un-profiled step)"}


def secondary_metrics(ssnamg, drv, state, m, n):
    import scipy.sparse as sp
    import torch
    ev = ssnamg.prox_residual(state["wk"], state["lk"], state["p"], state["q"], state["tk"], float("inf"), want=("Axprox", "s"))
    H0 = ssnamg.ASAt(ev["s"], state["p"], state["q"])
    Fk = state["bk1"] * state["lk"] - ev["Axprox"] - state["wlk"]
    # inner_solver = 2: PCG.m on Jk = bk1*I + H0/tk  (Class1/APD_SsN_Class1.m:149-152)
    Jk = (state["bk1"] * sp.identity(m + n, format="csr") + H0.to_scipy() / state["tk"]).tocsr()
    Jd = ssnamg.DeviceCSR.from_scipy(Jk)
    opts = {"retol": 1e-11, "maxit": 2000, "precd": 2, "guess": None}
    ssnamg.PCG(Jd, -Fk, opts)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    d, it, res, _ = ssnamg.PCG(Jd, -Fk, opts)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    pd = {"bk1": state["bk1"], "tk": state["tk"], "q": state["q"], "p": state["p"], "T": None, "H0": H0, "z": -Fk}
    ssnamg.rng_reset(); ssnamg.Hybrid_AMG(pd, drv.CLASS1_AMG_OPTIONS)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ssnamg.rng_reset(); zeta, itamg, resamg, info = ssnamg.Hybrid_AMG(pd, drv.CLASS1_AMG_OPTIONS)
    torch.cuda.synchronize(); dta = time.perf_counter() - t0
    # inner_solver = 5: Hybrid_twogrid on the same system (its iteration loop, coarse PCG included, is one cluster kernel)
    ssnamg.rng_reset(); ssnamg.Hybrid_twogrid(pd, drv.CLASS1_AMG_OPTIONS)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ssnamg.rng_reset(); _, ittg, restg, _ = ssnamg.Hybrid_twogrid(pd, drv.CLASS1_AMG_OPTIONS)
    torch.cuda.synchronize(); dtt = time.perf_counter() - t0
    return {"pcg_iters_per_s": it / dt, "pcg_iters": it, "pcg_rel_res": res, "pcg_system": f"Jk {m + n}x{m + n}, nnz {Jk.nnz}",
            "hybrid_amg_ms": dta * 1e3, "wcycles_per_s_incl_setup": itamg / dta,
            "hybrid_twogrid_ms": dtt * 1e3, "twogrid_iterations": int(ittg), "twogrid_rel_res": float(restg)}


METRIC2 = "ssn_amg_inner_solve_step_time_class2_64x64_grid"


def class2_config(info):
    return {"workload": "partial_OT_grid64x64_vs_64x64_m4096_n4096_outer1_ssn2_from_trivial_start", "E_active": int(info["E"]),
            "nnz_H0": int(info["nnzH"]), "amg_cycles": int(info["itamg"]), "line_search_trials": int(info["ll"]) + 1,
            "l2_flush": "inputs larger than L2 (wk and phi: 2 x 134 MB per pass)"}


def run_class2(args):
    """--config class2_64 (BASELINE configs[2]): one SsN step of Class2/APD_SsN_Class2.m:137-217 on the 64x64 grids (m = n =
    4096, 16.8M-entry plan, phi = 1, mu = 0.65 of the mass) at outer iteration 1, SsN step 2 from the trivial start: fused
    residual of partial OT (ssn_prox_residual_pot) -> ASAt -> AMG4POT -> line search -> new residual."""
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(0)
    import ssnamg
    drv = ssnamg.driver
    peak, peak_src = peaks()
    P = ssnamg.problems.grid_problem_pot(64, seed=0)
    m = n = 64 * 64
    ssnamg.rng_reset()
    st = drv.class2_trivial_state(P)
    st["lk"], _, _ = drv.ssn_step_class2(st)                 # SsN step 1 (untimed): the state of step 2

    def step_fn():
        ssnamg.rng_reset()
        return drv.ssn_step_class2(st)
    for _ in range(max(args.warmup, 3)):
        lk_new, Fk_new, info = step_fn()
    torch.cuda.synchronize()
    sampler = ClockSampler(0); sampler.start()
    l0 = ssnamg.launch_count()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        lk_new, Fk_new, info = step_fn()
    e1.record(); torch.cuda.synchronize()
    launches = ssnamg.launch_count() - l0
    ms_step = e0.elapsed_time(e1) / args.steps
    # the plan-wide kernel of this path, timed alone (CUDA events around the launch, ssn_kernel_timer)
    call = lambda: ssnamg.prox_residual_pot(st["wk"], st["lk"], st["p"], st["q"], st["tk"], st["phi"], want=("Hprox", "s", "t"))
    for _ in range(3):
        call()
    ssnamg.kernel_timer(True)
    for _ in range(20):
        call()
    kms, kcnt = ssnamg.kernel_timer_read(); ssnamg.kernel_timer(False)
    k_ms = kms / max(kcnt, 1)
    sampler.stop_flag = True; sampler.join(timeout=2)
    ops_runs = []
    for _ in range(7):                                       # the first one pays the allocations of the operator path: median of the other six
        ssnamg.rng_reset()
        _, _, oi = drv.ssn_step_class2_ops(st)
        ops_runs.append(oi)
    print("bench: operator-level step, amg4pot ms per run: " + ", ".join(f"{o['ms_amg']:.1f}" for o in ops_runs), file=sys.stderr)
    ops_info = sorted(ops_runs[1:], key=lambda o: o["ms_amg"])[len(ops_runs[1:]) // 2]
    bytes_pass = 16.0 * m * n + 1.0 * m * n                  # wk and phi read, s written
    ach = bytes_pass / (k_ms * 1e-3) / 1e9
    out = {"metric": METRIC2, "value": ms_step, "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "impl": "ours", "config": class2_config(info),
           "breakdown_ms": {"plan_wide_kernels": ops_info["ms_plan"], "asat_assembly": ops_info["ms_asat"], "amg4pot": ops_info["ms_amg"],
                            "note": "host-timed phases of the same step run operator by operator (each closed by a device synchronise); "
                                    "the timed steps are ONE library call each (ssn_ssn_step_class2)"},
           "gpu_launches": int(launches), "clocks": sampler.summary(),
           "roofline": {"bound": "hbm", "kernel": "plan_reduce_kernel<PROX, G_PHI> (fused residual of partial OT: z, prox, H*prox, ||prox||^2, "
                                                   "active flags; one read of wk and one of phi)",
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": bytes_pass, "avg_launch_ms": k_ms, "launches_per_step": int(info["ll"]) + 2,
                        "share_of_step": (int(info["ll"]) + 2) * k_ms / ms_step, "traffic": None,
                        "note": "285 MB per launch: the kernel lasts ~50 us, so launch ramp and tail weigh more than at 128x128"}}
    # e2e: the same step as ONE plugin call on HOST buffers (ssn_ssn_step_class2_host): wk, phi, the duals and the weights
    # are copied from pinned host memory inside the call every step, lk_new and Fk_new are copied back
    host = {k: (v.cpu().pin_memory() if isinstance(v, torch.Tensor) else v) for k, v in st.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values() if isinstance(v, torch.Tensor))

    def e2e_step():
        ssnamg.rng_reset()
        a, b, _ = drv.ssn_step_class2(host, host_call=True)
        return a, b
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ke = max(3, min(args.steps, 5))
    for _ in range(ke):
        lk_h, Fk_h = e2e_step()
    torch.cuda.synchronize()
    out["e2e"] = {"value": (time.perf_counter() - t0) * 1e3 / ke, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                  "d2h_bytes_per_step": int(lk_h.numel() * 8 + Fk_h.numel() * 8)}
    if not args.no_cpu_baseline:
        from oracle import bench_step
        Pn = load_problems_module().grid_problem_pot(64, seed=0)
        cms, lk_c, Fk_c, cinfo, t_state = bench_step.timed_step_class2(Pn)
        out["cpu_baseline"] = {"value": cms, "unit": UNIT, "cores": 1, "kind": "port",
                               "sample": f"oracle (NumPy/SciPy port of the reference), the WHOLE step once: residual {cinfo['phases_s']['residual_s']:.2f} s, "
                                         f"ASAt {cinfo['phases_s']['asat_s']:.2f} s, AMG4POT {cinfo['phases_s']['amg4pot_s']:.2f} s, line search "
                                         f"{cinfo['phases_s']['line_search_s']:.2f} s (whole-vector expressions; state built by {t_state:.1f} s of SsN step 1)",
                               "same_step_as_device": {"config_equal": class2_config(cinfo) == out["config"],
                                                       "lk_new_max_rel_diff": float(np.max(np.abs(lk_c - lk_new.cpu().numpy())) / np.max(np.abs(lk_c)))}}
    emit(out)
    return 0


def run_reference_class2(args):
    """--impl reference --config class2_64: the oracle's step on the host, never importing torch or the product."""
    from oracle import bench_step
    P = load_problems_module().grid_problem_pot(64, seed=0)
    ms, lk_new, Fk_new, info, t_state = bench_step.timed_step_class2(P)
    emit({"metric": METRIC2, "value": ms, "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0, "ms_per_step": ms,
          "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
          "config": class2_config(info), "run": {"state_build_s": round(t_state, 1), "requested_steps": args.steps},
          "cpu_baseline": {"value": ms, "unit": UNIT, "cores": 1, "kind": "port", "sample": "the whole step once, nothing sampled"},
          "e2e": {"value": ms, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    return 0


def run_reference(args):
    """--impl reference: the reference's CPU path (the oracle port; MATLAB/Octave are absent) on the host cores --
    the SAME step at full size, built from the committed state fixture and the problem generator only.  This
    process never imports torch or the product package and never loads libssnamg.so; nothing is sampled or
    extrapolated, so it runs ONE step whatever --steps / --warmup say, and says so."""
    g = args.grid
    m = n = g * g
    fixture = FIXTURE.format(g=g, k=args.state_outer)
    if not os.path.exists(fixture):
        emit({"impl": "reference", "unavailable": f"state fixture {os.path.relpath(fixture, ROOT)} missing (tools/save_bench_state.py writes it on a B200)"})
        return 0
    from oracle import bench_step
    t0 = time.time()
    P = load_problems_module().grid_problem(g, seed=0)
    st = bench_step.state_from_fixture(fixture, P)
    del P
    t_state = time.time() - t0
    workload = f"grid{g}x{g}_vs_{g}x{g}_m{m}_n{n}_outer{st['k']}_ssn{st['ssn_it']}"
    ms, lk_new, Fk_new, info, txt = cpu_step_full(st)
    ex = st["expect"]
    check = {"device_step_recorded_in_fixture": {k: (int(v) if np.ndim(v) == 0 and float(v).is_integer() else float(v)) for k, v in ex.items() if np.ndim(v) == 0}}
    if "lk_new" in ex:
        check["lk_new_max_rel_diff_vs_device"] = float(np.max(np.abs(lk_new - ex["lk_new"])) / np.max(np.abs(ex["lk_new"])))
    if "wlk" in ex:
        check["wlk_max_rel_diff_vs_device"] = float(np.max(np.abs(st["wlk"] - ex["wlk"])) / np.max(np.abs(ex["wlk"])))
    out = {"metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
           "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic", "impl": "reference", "config": step_config(workload, info, info["ll"]),
           "run": {"state_build_s": round(t_state, 1), "requested_steps": args.steps, "requested_warmup": args.warmup,
                   "note": "one full-size step is about two minutes of host time: it is run ONCE, not sampled and scaled"},
           "cpu_baseline": {"value": ms, "unit": UNIT, "cores": info["threads"], "kind": "port", "sample": txt},
           "e2e": {"value": ms, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "parity_with_device_step": check}
    emit(out)
    return 0


if __name__ == "__main__":
    sys.exit(main())
