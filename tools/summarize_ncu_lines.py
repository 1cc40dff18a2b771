#!/usr/bin/env python
"""Warp-stall samples per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`
(the library is built with -lineinfo): the 45 hottest lines of every captured launch.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass | python tools/summarize_ncu_lines.py
"""
import csv
import sys


def main():
    rows = list(csv.reader(sys.stdin))
    i, launch = 0, -1
    while i < len(rows):
        r = rows[i]
        if r and r[0] == "Line No":
            hdr = r
            si = hdr.index("# Samples")
            ei = hdr.index("Instructions Executed")
            launch += 1
            per = {}
            cur = None
            i += 1
            while i < len(rows) and rows[i] and rows[i][0] not in ("File Path", "Line No", "Function Name", "Kernel Name"):
                q = rows[i]
                if q[0] != "":
                    cur = (int(q[0]), q[1].strip())
                    per.setdefault(cur, [0, 0, 0])
                elif cur is not None and len(q) > max(si, ei):
                    try:
                        per[cur][0] += int(q[si])
                        per[cur][1] += int(q[ei])
                        per[cur][2] += 1
                    except ValueError:
                        pass
                i += 1
            tot = sum(v[0] for v in per.values()) or 1
            print(f"== source block {launch}: {tot} samples")
            for (ln, src), v in sorted(per.items(), key=lambda kv: -kv[1][0])[:45]:
                print(f"  {v[0]:6d} {100.0 * v[0] / tot:5.1f}%  sass {v[2]:5d} exec {v[1]:9d}  L{ln:<4d} {src[:110]}")
            continue
        i += 1


if __name__ == "__main__":
    main()
