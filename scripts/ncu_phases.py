"""Cuts the per-instruction counters of an ncu report of k_task_panda_lanes (built with -DB2_LANES_MARKERS) into the
phases between its PMTRIG markers: executed warp instructions, shared-memory wavefronts, stall samples and the share of
fp64 instructions per phase.  usage: python scripts/ncu_phases.py report.ncu-rep [warps]"""
import csv
import subprocess
import sys

NAMES = {None: "prologue (staging, loads)", 0: "PID", 1: "joint placement (sincos)", 2: "forward kinematics", 3: "C: S, inertia",
         4: "D: V prefix, Ic suffix", 5: "E: cdq, F = Ic S", 6: "F: acceleration prefix", 7: "G: body force, M rows",
         8: "H/I: force suffix, rhs", 9: "J: LDL^T", 10: "substitutions, integrate, rows check", 11: "obs: joint placement",
         12: "obs: forward kinematics", 13: "obs: assembly, reward", 14: "obs: copy out", 15: "state / PID store"}
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[1]
col = {k: hdr.index(k) for k in ("Source", "Instructions Executed", "# Samples", "L1 Wavefronts Shared", "Thread Instructions Executed")}
nw = float(sys.argv[2]) if len(sys.argv) > 2 else 5464.0
cur, acc, order = None, {}, []
for r in rows[2:]:
    try:
        inst, samp, wf, thr = (int(r[col[k]]) for k in ("Instructions Executed", "# Samples", "L1 Wavefronts Shared",
                                                       "Thread Instructions Executed"))
    except (ValueError, IndexError):
        continue
    text = r[col["Source"]]
    if "PMTRIG" in text:
        cur = int(text.split("PMTRIG")[1].strip().rstrip(";").strip(), 0)
        cur = cur.bit_length() - 1 if cur > 0 else 0    # PMTRIG takes a mask: bit n = event n
        continue
    if cur not in acc:
        acc[cur] = [0, 0, 0, 0, 0]
        order.append(cur)
    a = acc[cur]
    a[0] += inst; a[1] += samp; a[2] += wf; a[3] += thr
    if any(op in text for op in ("DFMA", "DMUL", "DADD")):
        a[4] += inst
ti, ts, tw = (sum(a[k] for a in acc.values()) or 1 for k in (0, 1, 2))
print(f"{'phase':42s} {'instr/warp':>10s} {'%':>5s} {'fp64/warp':>9s} {'smem wf/warp':>12s} {'%':>5s} {'samples %':>9s} {'lanes':>5s}")
for k in order:
    a = acc[k]
    if a[0] == 0:
        continue
    print(f"{NAMES.get(k, str(k)):42s} {a[0] / nw:10.1f} {100 * a[0] / ti:5.1f} {a[4] / nw:9.1f} {a[2] / nw:12.1f} {100 * a[2] / tw:5.1f} "
          f"{100 * a[1] / ts:9.1f} {a[3] / a[0]:5.1f}")
print(f"{'total':42s} {ti / nw:10.1f} {'':5s} {sum(a[4] for a in acc.values()) / nw:9.1f} {tw / nw:12.1f}")
