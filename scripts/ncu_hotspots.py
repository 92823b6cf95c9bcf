"""Source-level hot spots of one kernel in an ncu report captured with --import-source on (compile with -lineinfo):
warp-stall samples and executed warp instructions per source file and per source line, largest first.
usage: python scripts/ncu_hotspots.py report.ncu-rep [top_lines]"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, lines, files = None, {}, {}
for r in csv.reader(raw.splitlines()):
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        cur = r[1].split("/")[-1]
        continue
    if len(r) < 8 or r[0] in ("Line No", ""):
        continue
    try:
        ln, inst, samp = int(r[0]), int(r[7]), int(r[6])
    except ValueError:
        continue
    x = lines.setdefault((cur, ln), [0, 0, r[1][:110]])
    x[0] += inst
    x[1] += samp
    f = files.setdefault(cur, [0, 0])
    f[0] += inst
    f[1] += samp
total = sum(v[1] for v in files.values()) or 1
print(f"total samples {total}, warp instructions {sum(v[0] for v in files.values())}")
for k, v in sorted(files.items(), key=lambda kv: -kv[1][1]):
    print(f"file {k:24s} samples {v[1]:7d} ({100 * v[1] / total:5.1f} %)  instructions {v[0]}")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{k[0][:18]:18s} {k[1]:5d} samples {v[1]:6d} ({100 * v[1] / total:4.1f} %) inst {v[0]:9d} | {v[2]}")
