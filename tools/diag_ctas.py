"""Start/end times of every CTA of the last layer-wise tcgen05 GEMM launch of the SDF normals pass (debug)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vdn_nerf_b200 import configs, fields, ops, _lib
dev = "cuda"
conf = configs.CONFIGS["womsk_white"]
mods = configs.build_networks(conf, fields, seed=0, device=dev)
sdf = mods[1]
ops.set_precision("tf32")
lib = _lib.load()
n = 148 * 2 * 128 * 2
x = (torch.rand(n, 3, device=dev) * 2 - 1).requires_grad_(True)
for _ in range(2):
    sdf(x); sdf.gradient(x)
buf = torch.zeros(8192, dtype=torch.int64, device=dev)
torch.cuda.synchronize()
sdf(x)
lib.vdn_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
sdf.gradient(x)
torch.cuda.synchronize()
lib.vdn_debug_timeline(None)
t = buf.cpu().tolist()
nct = n // 128
rows = [(t[1024 + 4 * i], t[1024 + 4 * i + 1], t[1024 + 4 * i + 2]) for i in range(nct)]
t0 = min(r[0] for r in rows)
durs = [r[1] - r[0] for r in rows]
print("CTAs", nct, "kernel span %.1f us" % ((max(r[1] for r in rows) - t0) / 1e3))
print("duration us: min %.1f median %.1f max %.1f" % (min(durs) / 1e3, sorted(durs)[nct // 2] / 1e3, max(durs) / 1e3))
for i in list(range(0, 8)) + list(range(296, 304)) + list(range(nct - 4, nct)):
    r = rows[i]
    print(f"cta {i:4d} sm {r[2]:3d} start {(r[0]-t0)/1e3:7.1f} us dur {(r[1]-r[0])/1e3:6.1f} us")
r = lambda role, ev: t[role * 64 + ev] - t[0] if t[role * 64 + ev] else None
print("CTA0: alloc+sync done", r(0, 1), " acc ready", r(0, 2), " epilogue done", r(0, 3), " dealloc", r(0, 4))
for kb in range(8):
    print(f"  kb{kb}: prod loads-ready {r(1,3*kb)} empty-ok {r(1,3*kb+1)} arrived {r(1,3*kb+2)} | mma wait {r(2,2*kb)} full-ok {r(2,2*kb+1)}")
