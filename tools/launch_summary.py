"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X.csv ...`).
usage: python tools/launch_summary.py gpurun_out/X.csv "title" > profiles/<name>_summary.md"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = OrderedDict()
for r in rows[1:]:
    name = r[ki]
    if "awm::" in name:
        name = re.sub(r"\(.*", "", name).replace("void ", "")
    else:
        name = "(torch: synthetic input generation / conversion / copies)"
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += float(r[vi].replace(",", "")) / 1e6
allms = sum(t[1] for t in tot.values())
print("# %s\n" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("`ncu --metrics gpu__time_duration.sum --clock-control none`; per-launch times are serialised and cold-cache: only the SHARE of a kernel is\n"
      "comparable with the CUDA-event table of the bench line.  %d launches captured, %.3f ms in total.\n" % (len(rows) - 1, allms))
print("| kernel | launches | total ms | ms per launch | share |\n|---|---|---|---|---|")
for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.3f | %.4f | %.1f %% |" % (name, n, ms, ms / n, 100 * ms / allms))
