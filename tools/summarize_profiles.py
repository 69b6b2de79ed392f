"""Turns the ncu outputs a gpurun call left in gpurun_out/ into the committed summaries under profiles/:
  * launches_<tag>.csv  (ncu --metrics gpu__time_duration.sum launch list of one bench.py step)
      -> profiles/launch_summary_<tag>.md  (per-kernel count / total / share of the step)
  * <name>_full_<tag>.ncu-rep (ncu --set full captures)
      -> profiles/<name>_full_<tag>.csv (selected raw metrics) and profiles/traffic.json
Usage: python tools/summarize_profiles.py <tag>     (e.g. r1)
"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def to_bytes(v, unit):
    f = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit)
    return float(v.replace(",", "")) * f if f else None


def launch_summary(tag):
    path = os.path.join(GP, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    r = csv.reader(lines); hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict(); tot = 0.0; n = 0
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", "")); u = row[ui]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)          # -> us
        name = re.sub(r"\(.*", "", row[ki]).replace("void ", "").replace("ssn::<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; n += 1
    with open(os.path.join(OUT, f"launch_summary_{tag}.md"), "w") as f:
        f.write(f"# Launch list of one bench.py step ({tag})\n\n"
                "`SSN_BENCH_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv "
                "python bench.py --steps 1 --warmup 3 --no-cpu-baseline` (cold-cache, serialised per-launch times: shares, "
                "not absolutes, are comparable with the live run).\n\n"
                f"{n} launches, {tot / 1e3:.2f} ms of kernel time.\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k[:90]}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |\n")
    print("wrote launch summary:", n, "launches", round(tot / 1e3, 2), "ms")


def full_capture(name, tag, traffic):
    rep = os.path.join(GP, f"{name}_full_{tag}.ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    keep = [hdr.index("Kernel Name")] + [hdr.index(m) for m in METRICS if m in hdr]
    with open(os.path.join(OUT, f"{name}_full_{tag}.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in keep]); w.writerow([units[i] for i in keep])
        for r in rows[2:]:
            w.writerow([r[i] for i in keep])
    ir, iw, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    vals = [(to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]), r[hdr.index("Kernel Name")]) for r in rows[2:]]
    warm = vals[1:] if len(vals) > 1 else vals                               # the first capture also pays first-touch writes
    traffic[name] = {"dram_bytes_per_launch": sum(v for v, _ in warm) / len(warm), "kernel": warm[0][1][:80],
                     "captures": len(vals), "source": f"profiles/{name}_full_{tag}.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"}
    print("wrote", name, traffic[name]["dram_bytes_per_launch"])


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(OUT, exist_ok=True)
    launch_summary(tag)
    tpath = os.path.join(OUT, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    for name in ("k3", "trials", "screen", "spgemm", "cycle"):
        full_capture(name, tag, traffic)
    json.dump(traffic, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    main()
