"""Device time of the SDF network passes on 65536 points (training-step shape), tensor-core mode."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vdn_nerf_b200 import configs, fields, ops
dev = "cuda"
conf = configs.CONFIGS["womsk_white"]
mods = configs.build_networks(conf, fields, seed=0, device=dev)
sdf = mods[1]
ops.set_precision(sys.argv[1] if len(sys.argv) > 1 else "tf32")
n = 65536
x = (torch.rand(n, 3, device=dev) * 2 - 1).requires_grad_(True)

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

def value_only():
    with torch.no_grad():
        sdf.sdf(x)

def fwd():
    return sdf(x)

def fwd_normals():
    out = sdf(x)
    g = sdf.gradient(x)
    return out, g

def full():
    out = sdf(x)
    g = sdf.gradient(x)
    loss = out.sum() + (g * g).sum()
    loss.backward()

print("value-only %.3f ms" % timeit(value_only))
print("forward (saving) %.3f ms" % timeit(fwd))
print("forward + normals %.3f ms" % timeit(fwd_normals))
print("forward + normals + backward %.3f ms" % timeit(full))
