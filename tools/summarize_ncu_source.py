#!/usr/bin/env python
"""Condense `ncu -i X.ncu-rep --page source --csv --print-source sass` into a short text: for every captured launch the
warp-stall samples by SASS opcode and the 25 instructions with the most samples (with executed counts).

usage: ncu -i X.ncu-rep --page source --csv --print-source sass | python tools/summarize_ncu_source.py > profiles/r02_ncu_source_X.txt
"""
import collections
import csv
import re
import sys


def main():
    rows = list(csv.reader(sys.stdin))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for n, a in enumerate(starts):
        b = starts[n + 1] if n + 1 < len(starts) else len(rows)
        hdr = rows[a + 1]
        ci = {h: i for i, h in enumerate(hdr)}
        body = [r for r in rows[a + 2:b] if len(r) > 5 and r[0].startswith("0x")]
        if "# Samples" not in ci:
            continue
        tot = sum(int(r[ci["# Samples"]] or 0) for r in body) or 1
        print(f"== launch {n}: {rows[a][1]}   SASS instructions {len(body)}, stall samples {tot}")
        ops = collections.Counter()
        for r in body:
            m = re.sub(r"@!?U?P\d+\s+", "", r[1].strip()).split()[0]
            ops[m.split(".")[0]] += int(r[ci["# Samples"]] or 0)
        print("  samples by opcode: " + ", ".join(f"{k} {100.0 * v / tot:.1f}%" for k, v in ops.most_common(14)))
        top = sorted(body, key=lambda r: -int(r[ci["# Samples"]] or 0))[:25]
        for r in top:
            s = int(r[ci["# Samples"]] or 0)
            print(f"  {s:6d} {100.0 * s / tot:5.1f}%  executed {r[ci['Instructions Executed']]:>9s}  {r[1].strip()[:100]}")


if __name__ == "__main__":
    main()
