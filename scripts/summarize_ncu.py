"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_ppo_eager_r1.csv profiles/launches_ppo_step_r1.md
    python scripts/summarize_ncu.py full gpurun_out/gae_prof_r1.ncu-rep profiles/gae_scan_ncu_r1 [kernel-regex]
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]


def short(n):
    n = re.sub(r"<.*", "", n).replace("void ", "")
    return n[:72]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    rows = []
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        v = v / 1000 if row[ui] == "ns" else (v * 1000 if row[ui] == "ms" else v)
        rows.append((row[ki], v))
    names = [n for n, _ in rows]
    g0 = next(i for i, n in enumerate(names) if "gae_chunked" in n)
    g1 = next(i for i, n in enumerate(names) if "row_copy" in n and i > g0)
    phases = [("rollout: 8 collector steps", rows[:g0]), ("epoch prologue: GAE + cached old log-probs", rows[g0:g1]),
              ("update: 4 minibatches", rows[g1:])]
    out = ["# ncu launch list of the PPO step (config 2: N=4096, T=128, MLP(256,256), minibatch 16384)", "",
           "Source: `ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off` over "
           "`scripts/profile_step.py` (eager path so that every kernel is listed; 8 collector steps + GAE/old-logp "
           "+ 4 minibatches).  Times are cold-cache and serialised by ncu: compare SHARES, not absolutes.", "",
           "Total: %d launches, %.1f us.  `trl::*` are this repo's kernels; everything else is PyTorch/cuBLAS "
           "(the policy/value MLPs, which the north star keeps in PyTorch)." % (len(rows), sum(v for _, v in rows)), ""]
    for title, rs in phases:
        tot = sum(v for _, v in rs)
        mine = sum(v for n, v in rs if "trl::" in n)
        out += ["## %s -- %d launches, %.1f us (trl:: kernels %.1f us = %.1f%%)" % (title, len(rs), tot, mine,
                                                                                 100 * mine / tot), "",
                "| kernel | launches | total us | share | avg us |", "|---|---:|---:|---:|---:|"]
        agg = collections.OrderedDict()
        for n, v in rs:
            a = agg.setdefault(short(n), [0, 0.0])
            a[0] += 1
            a[1] += v
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.append("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, c, t, 100 * t / tot, t / c))
        out.append("")
    open(dst, "w").write("\n".join(out))
    print("wrote", dst)


def full(rep, dst, regex=None):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    res = []
    for row in data:
        if regex and not re.search(regex, row[ki]):
            continue
        d = {"kernel": row[ki]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                try:
                    d[k] = float(row[i].replace(",", ""))
                except ValueError:
                    d[k] = row[i]
                d[k + "__unit"] = units[i]
        res.append(d)
    json.dump(res, open(dst + ".json", "w"), indent=1)
    print("wrote", dst + ".json", len(res), "launches")
    return res


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)
