"""Timeline of CTA 0 of a tcgen05 GEMM launch (debug)."""
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
for n in (128, 148 * 2 * 128, 1 << 21):
    x = torch.rand(n, 3, device=dev) * 2 - 1
    for _ in range(2):
        sdf(x)
    buf = torch.zeros(8192, dtype=torch.int64, device=dev)
    lib.vdn_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
    torch.cuda.synchronize()
    sdf(x)
    torch.cuda.synchronize()
    lib.vdn_debug_timeline(None)
    t = buf.cpu().tolist()
    t0 = t[0]
    r = lambda role, ev: t[role * 64 + ev] - t0 if t[role * 64 + ev] else None
    print(f"--- n={n} (last launch of the chain; cycles since CTA start)")
    print("  alloc+sync done", r(0, 1), " acc ready", r(0, 2), " epilogue done", r(0, 3), " dealloc", r(0, 4))
    for kb in range(8):
        print(f"  kb{kb}: producer wait-empty {r(1,3*kb)} empty-ok/cp.async issued {r(1,3*kb+1)} finished+arrived {r(1,3*kb+2)} | "
              f"mma wait {r(2,2*kb)} full-ok {r(2,2*kb+1)}")
