#!/usr/bin/env python3
"""Per-source-line view of an ncu --import-source capture: executed warp instructions, stall samples and L1 wavefronts
(shared + global tag requests) attributed to CUDA source lines.  Usage: tools/source_lines.py <rep> [top]"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
agg = collections.OrderedDict()
cur_file, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    g = dict(zip(hdr, r))
    def num(k):
        try:
            return float(g.get(k, "0").replace(",", ""))
        except ValueError:
            return 0.0
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [r[1].strip(), 0.0, 0.0, 0.0, 0.0])
    a[1] += num("Instructions Executed")
    a[2] += num("# Samples")
    a[3] += num("L1 Wavefronts Shared")
    a[4] += num("L1 Tag Requests Global")
ti = sum(v[1] for v in agg.values()); ts = sum(v[2] for v in agg.values()); tw = sum(v[3] + v[4] for v in agg.values())
print("total warp instructions %.0f, stall samples %.0f, L1 wavefronts (shared) + tag requests (global) %.0f" % (ti, ts, tw))
byf = collections.defaultdict(lambda: [0.0, 0.0, 0.0])
for (f, _), v in agg.items():
    byf[f][0] += v[1]; byf[f][1] += v[2]; byf[f][2] += v[3] + v[4]
for f, v in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print("  %-20s %5.1f %% instr  %5.1f %% stalls  %5.1f %% L1 traffic" % (f, 100 * v[0] / ti, 100 * v[1] / ts, 100 * v[2] / max(tw, 1)))
print("%-18s %5s %7s %7s %7s  %s" % ("file", "line", "instr%", "stall%", "L1%", "source"))
for (f, ln), v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-18s %5d %6.2f%% %6.2f%% %6.2f%%  %s" % (f, ln, 100 * v[1] / ti, 100 * v[2] / ts, 100 * (v[3] + v[4]) / max(tw, 1), v[0][:110]))
