#!/usr/bin/env python
"""Turn the artefacts of one GPU-box visit (scripts/gpu_check.sh) into the text summaries kept
under profiles/: the per-kernel launch list shares, the headline ncu metrics of the captured
kernels, and (optionally) the opcode mix + warp-stall reasons from the ncu source page.

    python scripts/ncu_summary.py gpurun_out/<tag> profiles/<name>.md [--source REGEX]

Runs on the CPU box (ncu -i reads the report; no GPU needed).
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum", "smsp__inst_executed.sum",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_fma.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_elapsed.max",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def launch_shares(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ci = {n: i for i, n in enumerate(h)}
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(h) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        k = r[ci["Kernel Name"]].split("(")[0].replace("void ", "")
        v = float(r[ci["Metric Value"]].replace(",", ""))
        f = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ci["Metric Unit"]], 1.0)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v * f
    return agg


def raw_metrics(rep):
    out = ncu(["-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(out)))
    if not rows:
        return []
    h, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = collections.OrderedDict()
        d["kernel"] = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
        for m in RAW:
            if m in h:
                d[m] = r[h.index(m)] + " " + units[h.index(m)]
        res.append(d)
    return res


def source_mix(rep, regex):
    out = ncu(["-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + regex])
    rows = list(csv.reader(io.StringIO(out)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    if not hdr:
        return None
    h = rows[hdr[0]]
    body = rows[hdr[0] + 1:(hdr[1] - 1 if len(hdr) > 1 else len(rows))]
    ci = {n: i for i, n in enumerate(h)}
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = collections.Counter()
    opc = collections.Counter()
    ninst = samples = 0
    for r in body:
        if len(r) < len(h):
            continue
        n = int(r[ci["Instructions Executed"]])
        ninst += n
        op = r[ci["Source"]].split()
        o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        opc[o] += n
        samples += int(r[ci["# Samples"]])
        for s in stalls:
            tot[s] += int(r[ci[s]])
    return ninst, samples, opc, tot


def main():
    src, dst = sys.argv[1], sys.argv[2]
    regex = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else None
    particles = int(sys.argv[sys.argv.index("--particles") + 1]) if "--particles" in sys.argv else None
    L = ["# ncu summary of `%s`" % src, ""]
    for name in ("bench.json", "bench_ref.json"):
        p = os.path.join(src, name)
        if os.path.exists(p) and os.path.getsize(p):
            try:
                j = json.loads(open(p).read().strip().splitlines()[-1])
                L += ["## %s" % name, "", "```json", json.dumps(j, indent=1), "```", ""]
            except Exception:
                pass
    p = os.path.join(src, "launches.csv")
    if os.path.exists(p):
        agg = launch_shares(p)
        tot = sum(a[1] for a in agg.values())
        L += ["## launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised: "
              "compare shares, not absolutes)", "", "| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            L.append("| `%s` | %d | %.1f | %.1f %% |" % (k, a[0], a[1], 100 * a[1] / tot))
        L.append("")
    for f in sorted(os.listdir(src)):
        if not f.endswith(".ncu-rep"):
            continue
        rep = os.path.join(src, f)
        L += ["## `%s` (`ncu --set full --clock-control none --import-source on`)" % f, ""]
        for d in raw_metrics(rep):
            L.append("### %s" % d.pop("kernel"))
            L.append("")
            for k, v in d.items():
                L.append("- `%s` = %s" % (k, v))
            L.append("")
        if regex:
            sm = source_mix(rep, regex)
            if sm:
                ninst, samples, opc, tot = sm
                L += ["### opcode mix of `%s` (warp instructions executed: %d)" % (regex, ninst), ""]
                if particles:
                    L.append("per particle (x32 lanes / %d particles): %.0f thread-instructions" % (particles, ninst * 32.0 / particles))
                    L.append("")
                L += ["| opcode | warp instr | share |" + (" per particle |" if particles else ""),
                      "|---|---:|---:|" + ("---:|" if particles else "")]
                for k, v in opc.most_common(24):
                    L.append("| %s | %d | %.1f %% |" % (k, v, 100.0 * v / ninst) + (" %.1f |" % (v * 32.0 / particles) if particles else ""))
                L += ["", "### warp stall samples (%d)" % samples, "", "| reason | share |", "|---|---:|"]
                for k, v in tot.most_common(10):
                    L.append("| %s | %.1f %% |" % (k, 100.0 * v / max(samples, 1)))
                L.append("")
    os.makedirs(os.path.dirname(dst) or ".", exist_ok=True)
    open(dst, "w").write("\n".join(L) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
