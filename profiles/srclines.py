#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel from an ncu report.
  python profiles/srclines.py gpurun_out/prof.ncu-rep [kernel-regex] [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
kern = sys.argv[2] if len(sys.argv) > 2 else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kern: cmd += ["--kernel-name", "regex:" + kern]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None
lines = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] != "" and hdr:
        d = dict(zip(hdr, r))
        try:
            lines.append((cur_file, int(r[0]), r[1].strip(), int(d["Instructions Executed"]), int(d["# Samples"])))
        except ValueError:
            pass
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print("total warp instructions %d, stall samples %d" % (ti, ts))
print("%-16s %5s %7s %7s  %s" % ("file", "line", "inst%", "samp%", "source"))
for f, n, src, ins, smp in sorted(lines, key=lambda x: -x[3])[:top]:
    print("%-16s %5d %6.2f%% %6.2f%%  %s" % (f, n, 100.0 * ins / ti, 100.0 * smp / max(ts, 1), src[:90]))
