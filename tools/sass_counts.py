#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (B200_PROFILING.md): tcgen05.mma ->
UTC*MMA, tcgen05.ld / st -> LDTM / STTM, cp.async.bulk -> UBLKCP, cp.async.bulk.tensor -> UTMALDG, tcgen05.commit ->
UTCBAR, legacy mma.sync -> HMMA.

    python tools/sass_counts.py [vdn_nerf_b200/libvdn_b200.so] > profiles/r02_sass_counts.txt
"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "vdn_nerf_b200/libvdn_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WANT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "HMMA", "REDG",
        "MUFU.EX2", "SYNCS"]
counts = collections.OrderedDict()
cur = None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for w in WANT:
            if op.startswith(w):
                counts[cur][w] += 1
print("%-78s %7s " % ("kernel (mangled)", "instrs") + " ".join("%8s" % w for w in WANT))
for k, c in counts.items():
    if not any(c[w] for w in WANT[:9]):
        continue
    print("%-78s %7d " % (k[:78], c["_total"]) + " ".join("%8d" % c[w] for w in WANT))
