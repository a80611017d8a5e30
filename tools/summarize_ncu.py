#!/usr/bin/env python3
"""Turns ncu artefacts brought back in gpurun_out/ into the small text summaries committed under profiles/.
  launch list (csv of gpu__time_duration.sum)      -> per-kernel count / total / share / mean
  full report (.ncu-rep, via `ncu -i ... --page raw/source --csv`) -> key metrics + executed-opcode histogram"""
import collections, csv, json, subprocess, sys


def launch_summary(path):
    rows = list(csv.reader(open(path)))
    start = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[start]
    iK, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[start + 1:]:
        if len(r) <= iV:
            continue
        agg[r[iK].split("(")[0]][0] += 1
        agg[r[iK].split("(")[0]][1] += float(r[iV].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    out = ["kernel,launches,total_ms,share_pct,mean_us"]
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%s,%d,%.3f,%.2f,%.1f" % (k, n, v / 1e6, 100 * v / tot, v / n / 1e3))
    return "\n".join(out) + "\n"


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def report_summary(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:3]:
        g = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        out.append("kernel: %s  grid %s block %s" % (g.get("Kernel Name"), g.get("launch__grid_size"), g.get("launch__block_size")))
        for k in KEYS + [h for h in hdr if "issue_stalled" in h and "per_issue_active" in h]:
            if k in g:
                try:
                    if "issue_stalled" in k and float(g[k]) < 0.02:
                        continue
                except ValueError:
                    pass
                out.append("%s = %s %s" % (k, g[k], u[k]))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
    byop, tot = collections.Counter(), 0
    for r in rows[2:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        t = r[iS].split()
        op = t[1] if t[0].startswith("@") else t[0]
        byop[op.split(".")[0]] += int(r[iE])
        tot += int(r[iE])
    out.append("executed warp-instructions by opcode (share of %d static SASS lines' executions):" % (len(rows) - 2))
    for op, n in byop.most_common(28):
        out.append("  %-8s %6.2f %%" % (op, 100.0 * n / tot))
    return "\n".join(out) + "\n"


if __name__ == "__main__":
    kind, src, dst = sys.argv[1:4]
    open(dst, "w").write(launch_summary(src) if kind == "launches" else report_summary(src))
    print(open(dst).read())
