"""Timeline of CTA 0 of the fused SDF chain kernel (debug): per phase (slot, layer) the cycles at which the epilogue
warps and the MMA issuer pass their synchronisation points."""
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
n = 148 * 128 * 8
x = torch.rand(n, 3, device=dev) * 2 - 1
torch.set_grad_enabled(False)
for _ in range(2):
    sdf.sdf(x)
buf = torch.zeros(8192, dtype=torch.int64, device=dev)
lib.vdn_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
torch.cuda.synchronize()
sdf.sdf(x)
torch.cuda.synchronize()
lib.vdn_debug_timeline(None)
t = buf.cpu().tolist()
t0 = min(v for v in t if v)
g = lambda role, ph, ev: (t[role * 256 + 4 * ph + ev] - t0) if t[role * 256 + 4 * ph + ev] else -1
print("phase | epi: wait-start d_full-got drained a-written | mma: wait-start drained-ok a-ready-ok issued")
for ph in range(40):
    print(f"{ph:3d} | {g(0,ph,0):8d} {g(0,ph,1):8d} {g(0,ph,2):8d} {g(0,ph,3):8d} | {g(1,ph,0):8d} {g(1,ph,1):8d} {g(1,ph,2):8d} {g(1,ph,3):8d}")
