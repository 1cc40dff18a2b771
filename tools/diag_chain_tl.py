"""Timeline of one chain launch (debug build: make EXTRA_NVFLAGS=-DVDN_CHAIN_TL): per (tile pair, phase, slot) of CTA 0, what
the MMA issuer waited for and when each epilogue warp got its accumulator / finished.

    VDN_TL_LAUNCH=2 python tools/diag_chain_tl.py      # 0: SDF forward, 1: normals, 2: backward phase 1, 3: phase 2
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vdn_nerf_b200 import _lib, configs, fields, ops  # noqa: E402

dev = "cuda"
conf = configs.CONFIGS["womsk_white"]
mods = configs.build_networks(conf, fields, seed=0, device=dev)
sdf = mods[1]
ops.set_precision("tf32")
n = 65536
x = (torch.rand(n, 3, device=dev) * 2 - 1)
buf = torch.zeros(48 * 512, dtype=torch.int64, device=dev)
lib = _lib.load()
lib.vdn_debug_timeline(buf.data_ptr())
s, f, g = sdf.forward_split(x)
loss = s.sum() + (f * 0.01).sum() + (g * g).sum()
loss.backward()
torch.cuda.synchronize()
t = buf.cpu().view(-1, 48)
clk = 1.965e3      # cycles per microsecond
rows = [(i, r) for i, r in enumerate(t) if r[3] != 0]
t0 = min(int(r[0]) for _, r in rows)
print("it ph s |  issuer: start  drained  a_ready  issued | epilogue warps: full(min/max)  done(min/max) | wait of the fastest warp")
prev_done = {}
for i, r in rows:
    it, p, s_ = i // 32, (i // 2) % 16, i % 2
    u = lambda v: (int(v) - t0) / clk
    full = [int(r[4 + 2 * w]) for w in range(16)]
    done = [int(r[5 + 2 * w]) for w in range(16)]
    print(f"{it:2d} {p:2d} {s_} | {u(r[0]):7.2f} {u(r[1]):7.2f} {u(r[2]):7.2f} {u(r[3]):7.2f} | "
          f"{u(min(full)):7.2f} {u(max(full)):7.2f}  {u(min(done)):7.2f} {u(max(done)):7.2f} | "
          f"skew done {u(max(done)) - u(min(done)):5.2f}" +
          (" | warp 12 groups done: " + " ".join(f"{u(r[40 + k]):7.2f}" for k in range(8)) if int(r[40]) else ""))
