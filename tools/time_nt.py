"""Device time of the SDF normals pass (9 layer-wise tcgen05 GEMM launches) versus the number of 128-point tiles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vdn_nerf_b200 import configs, fields, ops
dev = "cuda"
conf = configs.CONFIGS["womsk_white"]
mods = configs.build_networks(conf, fields, seed=0, device=dev)
sdf = mods[1]
ops.set_precision("tf32")

def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

for waves in (0.25, 0.5, 1, 2, 4, 8):
    n = int(148 * 2 * 128 * waves)
    x = (torch.rand(n, 3, device=dev) * 2 - 1).requires_grad_(True)
    t_f = timeit(lambda: sdf(x))
    t_fn = timeit(lambda: (sdf(x), sdf.gradient(x)))
    print(f"n={n:8d} ({waves} waves of 296 CTAs): forward {t_f*1e3:8.1f} us, normals {1e3*(t_fn - t_f):8.1f} us = {1e3*(t_fn-t_f)/9:6.1f} us per launch")

# per-launch device time by CUDA events inside the library (prof hooks) vs wall/GPU time of the whole pass
import ctypes, time
from vdn_nerf_b200 import _lib
lib = _lib.load()
n = 148 * 2 * 128 * 2
x = (torch.rand(n, 3, device=dev) * 2 - 1).requires_grad_(True)
sdf(x); sdf.gradient(x); torch.cuda.synchronize()
lib.vdn_prof_enable(1)
sdf(x)
t0 = time.perf_counter()
sdf.gradient(x)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
ms, sp, fl = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
lib.vdn_prof_read(2, ctypes.byref(ms), ctypes.byref(sp), ctypes.byref(fl))
print(f"prof: {sp.value} layer-wise launches, {ms.value*1e3:.1f} us by events = {ms.value*1e3/max(1,sp.value):.1f} us each; "
      f"host time in gradient() {1e6*(t1-t0):.0f} us, until sync {1e6*(t2-t0):.0f} us")
lib.vdn_prof_enable(0)
