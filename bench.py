#!/usr/bin/env python
"""Benchmark of the neural-SDF volume-rendering hot path (contract: see the task statement / DESIGN.md).

    python bench.py --gpus N --steps K --warmup W                 # this repo's kernels
    python bench.py --impl reference --gpus N --steps K --warmup W  # the reference's CPU implementation (oracle port)

Workloads (BASELINE.json configs):
  train (default)  configs[1]: womsk_white.conf training step, 512 rays/batch/GPU, 64 coarse + 64 importance + 32
                   outside samples, render forward + driver loss + backward (+ NCCL gradient all-reduce for N > 1);
                   metric = train rays/s, whole job.  Weak scaling: every rank renders its own 512 rays.
  grid             configs[3]: extract_fields SDF query on the 512^3 grid, x-slabs sharded over the ranks;
                   metric = SDF pts/s.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic MAC / point (reverse-mode accounting, dense)
F_SDF, F_SDF1, G_SDF, F_COL, F_DEP, F_NERF, F_NERF_D = 524544, 459008, 458752, 271360, 295168, 604160, 616448


def alg_flops_per_ray(depth: bool) -> float:
    f_dep = F_DEP if depth else 0
    f_nerf = F_NERF_D if depth else F_NERF
    fwd = 112 * F_SDF1 + 128 * (F_SDF + G_SDF + F_COL + f_dep) + 160 * f_nerf
    bwd = 2 * (128 * (F_SDF + G_SDF + F_COL + f_dep) + 160 * f_nerf)
    return 2.0 * (fwd + bwd)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm": p["hbm_gbs"], "bf16_burst": p["bf16_tflops"], "bf16_sustained": p["bf16_tflops_sustained"],
                "source": "MEASURED_PEAKS.json"}
    return {"hbm": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
def setup_dist(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(local)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world, dev):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def ncu_traffic(name, rnd="r02"):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured launches) from the
    committed `ncu --set full` summary profiles/<rnd>_ncu_full_<name>.txt; None when the file is missing."""
    units = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    try:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "%s_ncu_full_%s.txt" % (rnd, name))
        tot, n = 0.0, 0
        for ln in open(path):
            f = ln.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * units.get(f[2], 1.0)
                n += f[0] == "dram__bytes_read.sum"
        return tot / n if n else None
    except Exception:   # noqa: BLE001 - a missing or malformed summary only means "no traffic figure"
        return None


def timed_steps(fn, steps, warmup, world, dev, flush, lib):
    """W untimed + K timed calls of fn(i); every timed call is bracketed by CUDA events on the current stream,
    an L2 flush (a 256 MiB write) runs untimed between calls.  Returns total milliseconds (max over ranks)."""
    for i in range(warmup):
        fn(i)
    barrier(world)
    evs = []
    l0 = lib.vdn_launch_count()
    align = torch.zeros(1, device=dev) if world > 1 else None
    for i in range(steps):
        flush.zero_()
        if align is not None:
            # untimed, like the flush: a one-element all-reduce lines the ranks up on the device, so that the skew the
            # flushes and the host introduce between steps is not billed to the first collective of the timed step
            import torch.distributed as dist
            dist.all_reduce(align)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn(warmup + i)
        b.record()
        evs.append((a, b))
    barrier(world)
    total = sum(a.elapsed_time(b) for a, b in evs)
    return max_over_ranks(total, world, dev), int(lib.vdn_launch_count() - l0)


def cpu_reference_step(n_rays, depth, reps, warm=1):
    """The reference algorithm (oracle port, PyTorch CPU, all host threads) on a bounded sample: rays/s."""
    from oracle import vdn_oracle as vo
    from vdn_nerf_b200 import configs, fields
    conf = configs.CONFIGS["womsk_white_wdepth" if depth else "womsk_white"]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    mods = configs.build_networks(conf, fields, seed=0)
    nets = vo.nets_from_modules(*mods, conf)
    for _, t in nets.leaves():
        t.requires_grad_(True)
    o, d, near, far = vo.synthetic_rays(n_rays)
    rgb = torch.full((n_rays, 3), 0.5)
    times = []
    for i in range(warm + reps):
        torch.manual_seed(2)
        t0 = time.perf_counter()
        out = vo.render(nets, o, d, near, far, background_rgb=torch.ones(1, 3), cos_anneal_ratio=1.0)
        gt = torch.full_like(out["render_feats"], 0.5) if out["render_feats"] is not None else None
        loss = vo.driver_loss(out, rgb, gt_feats=gt)
        torch.autograd.grad(loss, [t for _, t in nets.leaves()], allow_unused=True)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return n_rays * len(times) / sum(times), threads, times


def cpu_reference_grid(n_points, reps):
    from oracle import vdn_oracle as vo
    from vdn_nerf_b200 import configs, fields
    conf = configs.CONFIGS["womsk_white"]
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    mods = configs.build_networks(conf, fields, seed=0)
    nets = vo.nets_from_modules(*mods, conf)
    pts = torch.rand(n_points, 3) * 2.02 - 1.01
    times = []
    with torch.no_grad():
        for i in range(1 + reps):
            t0 = time.perf_counter()
            vo.sdf_value(nets.sdf, pts, nets.sdf_spec)
            if i:
                times.append(time.perf_counter() - t0)
    return n_points * len(times) / sum(times), threads, times


# ------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores: the UNMODIFIED reference staged under
    oracle/_ref (oracle/stage_ref.py) when it travelled with the snapshot, else the oracle port.  Same rays, weights and
    config as the CUDA arm (512 rays per step); rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import stage_ref
    depth = args.workload == "train_wdepth"
    have_ref = stage_ref.available()
    threads = os.cpu_count() or 1
    if args.workload == "grid":
        sample = 4 * 64 ** 3
        if have_ref:
            from oracle import ref_runner
            from vdn_nerf_b200 import configs
            torch.set_num_threads(threads)
            mods, _, _ = ref_runner.build(configs.CONFIGS["womsk_white"], "cpu")
            pts = torch.rand(sample, 3) * 2.02 - 1.01
            times = []
            with torch.no_grad():
                for i in range(args.warmup + args.steps):
                    t0 = time.perf_counter()
                    for blk in pts.split(64 ** 3):           # extract_fields queries 64^3 blocks (renderer.py:10-30)
                        mods[1].sdf(blk)
                    if i >= args.warmup:
                        times.append(time.perf_counter() - t0)
            v = sample * len(times) / sum(times)
            kind = "reference"
        else:
            v, threads, times = cpu_reference_grid(sample, args.warmup + args.steps)
            kind = "port"
        value, unit, metric, ms = v, "pts/s", "sdf_grid_pts_per_s", 1e3 * sum(times) / len(times)
        cfg = {"workload": "extract_fields SDF grid query, womsk_white SDF net, x-slabs per rank", "resolution": 512,
               "mode": "reference, PyTorch CPU fp32"}
        sample_s = f"{sample} lattice points per step (4 blocks of 64^3 of the 512^3 grid), SDFNetwork.sdf, PyTorch CPU fp32"
    else:
        sample = args.rays
        if have_ref:
            from oracle import ref_runner
            r = ref_runner.run("cpu", sample, args.steps, args.warmup, depth, threads)
            v, times = r["rays_per_s"], [t * 1e-3 for t in r["times_ms"]]
            kind = "reference"
        else:
            v, threads, times = cpu_reference_step(sample, depth, args.steps, warm=args.warmup)
            kind = "port"
        value, unit, metric, ms = v, "rays/s", "train_rays_per_s", 1e3 * sum(times) / len(times)
        cfg = {"workload": "womsk_white%s training step (BASELINE configs[%d]): render fwd + driver loss + bwd"
                           % ("_wdepth" if depth else "", 2 if depth else 1), "rays_per_step_per_gpu": args.rays,
               "global_batch": args.rays, "n_samples": 64, "n_importance": 64, "n_outside": 32,
               "mode": "reference, PyTorch CPU fp32 autograd"}
        sample_s = f"{sample} rays per step (the full batch), PyTorch CPU fp32 autograd, {threads} threads"
    line = {"impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": kind, "sample": sample_s},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": ("the unmodified reference classes (dpt_models/{embedder,fields,renderer}.py staged byte for byte under "
                     "oracle/_ref by oracle/stage_ref.py) on the host cores" if kind == "reference" else
                     "oracle/_ref was not staged on this box: this is oracle/vdn_oracle.py, the restatement pinned "
                     "bit-exact against the live reference")}
    print(json.dumps(line))


def reference_cuda_eager(rays, depth, steps=10, warmup=3, timeout=600):
    """The reference's own eager CUDA path on this GPU: the staged reference classes under
    torch.set_default_tensor_type('torch.cuda.FloatTensor') (dpt_runner.py:744), in a subprocess because the switch is
    process-wide.  Returns the parsed JSON of oracle/ref_runner.py or {"unavailable": why}."""
    from oracle import stage_ref
    if not stage_ref.available():
        return {"unavailable": "oracle/_ref not staged on this box"}
    cmd = [sys.executable, "-m", "oracle.ref_runner", "--device", "cuda", "--rays", str(rays), "--steps", str(steps),
           "--warmup", str(warmup)] + (["--depth"] if depth else [])
    try:
        out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": "no result: " + out.stderr.strip()[-300:]}
    except Exception as ex:    # noqa: BLE001 - a failed side measurement must not fail the bench line
        return {"unavailable": repr(ex)[:300]}


def run_ours(args):
    from oracle import vdn_oracle as vo          # synthetic-ray protocol + cpu_baseline leg only
    from vdn_nerf_b200 import _lib, configs, dist as vdist, fields, ops
    from vdn_nerf_b200.renderer import NeuSRenderer, extract_fields_sdf
    from vdn_nerf_b200.training import train_step
    rank, world, local = setup_dist(args.gpus)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    ops.set_precision(args.precision)
    pk = peaks()
    depth = args.workload == "train_wdepth"
    pose = args.workload == "train_pose"
    conf = configs.CONFIGS["womsk_white_wdepth" if depth else "womsk_white"]
    mods = configs.build_networks(conf, fields, seed=0, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    line = {}

    if args.workload == "grid":
        res = args.resolution
        lo, hi = vdist.shard_range(res, rank, world)
        u = torch.empty(hi - lo, res, res, device=dev)
        bmin, bmax = [-1.01] * 3, [1.01] * 3

        def fn(i):
            extract_fields_sdf(mods[1], bmin, bmax, res, x_range=(lo, hi), out=u)
        with ClockSampler(local) as cs:
            ms, launches = timed_steps(fn, args.steps, args.warmup, world, dev, flush, lib)
        pts = float(res) ** 3
        value = pts * args.steps / (ms * 1e-3)
        # end to end: bounds from host, field back on the host (the reference's extract_fields returns numpy)
        host_u = torch.empty(hi - lo, res, res, pin_memory=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(max(1, args.steps // 2)):
            fn(i)
            host_u.copy_(u, non_blocking=False)
        torch.cuda.synchronize()
        e2e_t = (time.perf_counter() - t0) / max(1, args.steps // 2)
        e2e_t = max_over_ranks(e2e_t, world, dev)
        lib.vdn_prof_enable(1)
        fn(0)
        import ctypes
        fam_ms, fam_n = 0.0, 0
        for f in (0, 2, 3):
            msn, sp, fl = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double()
            lib.vdn_prof_read(f, ctypes.byref(msn), ctypes.byref(sp), ctypes.byref(fl))
            fam_ms += msn.value
            fam_n += sp.value
        lib.vdn_prof_enable(0)
        local_pts = float(hi - lo) * res * res
        ach = 2.0 * F_SDF1 * local_pts / (fam_ms * 1e-3) / 1e12
        kname = ("sdf_chain_tc_kernel (fused PE + 9-layer tcgen05 kind::f16 chain, fp16 operands / fp32 accumulate, 1 launch per slab)"
                 if args.precision == "tf32" else "gemm_nt_kernel (FFMA), 9 launches per slab")
        line.update({"metric": "sdf_grid_pts_per_s", "unit": "pts/s", "value": value, "ms_per_step": ms / args.steps,
                     "config": {"workload": "extract_fields SDF grid query, womsk_white SDF net, x-slabs per rank",
                                "resolution": res, "l2": "256 MiB flush between timed iterations", "mode": args.precision},
                     "e2e": {"value": pts / e2e_t, "unit": "pts/s", "h2d_bytes_per_step": 3 * res * 4,
                             "d2h_bytes_per_step": int(local_pts * 4)},
                     "roofline": {"bound": "tensor", "kernel": kname,
                                  "achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                                  "frac": ach / pk["bf16_sustained"],
                                  "traffic": ncu_traffic("chain", "r01") if args.precision == "tf32" else None,
                                  "traffic_note": "DRAM bytes per launch of a 2 M-point slab (ncu --set full, "
                                                  "profiles/r01_ncu_full_chain.txt); algorithmic: 8 MB written",
                                  "peak_source": pk["source"] + " bf16 sustained (kind::f16 runs at the bf16 rate)",
                                  "launch_ms": fam_ms / max(1, fam_n)}})
        if args.precision == "tf32":
            line["dtype"] = "f16"
        cpu_kind = "grid"
    else:
        B = args.rays if args.global_batch <= 0 else max(1, args.global_batch // world)
        rend = NeuSRenderer(*mods, **conf["neus_renderer"])
        params = [p for m in mods if m is not None for p in m.parameters()]
        o, d, near, far = (t.to(dev) for t in vo.synthetic_rays(B, seed=1234 + rank))
        rgb = torch.full((B, 3), 0.5, device=dev)
        gt = torch.full((B, 96), 0.5, device=dev) if depth else None
        bg = torch.ones(1, 3, device=dev)
        sync = vdist.FlatGradAllReduce(params) if world > 1 else None
        if pose:            # BASELINE configs[4]: gradients also reach the rays (stand-in for the so(3) pose refinement)
            o.requires_grad_(True)
            d.requires_grad_(True)

        def fn(i):
            torch.manual_seed(2 + i)
            if pose:
                o.grad = d.grad = None
            train_step(rend, params, o, d, near, far, rgb, gt_feats=gt, background_rgb=bg, cos_anneal_ratio=1.0,
                       grad_sync=sync, global_batch=B * world)
        gstep = None
        if not args.no_cuda_graph and (world == 1 or not args.no_graph_nccl):
            # the whole step (render + loss + backward) captured once, replayed per step: one graph launch instead of
            # ~250 kernel launches; the kernels and their work are unchanged.  Falls back to eager launches if the
            # capture is refused.  Multi-GPU steps are captured together with their two NCCL all-reduces
            # (capture_error_mode="thread_local", see training.GraphedTrainStep).
            from vdn_nerf_b200.training import GraphedTrainStep
            fn(0)                                       # eager warm-up: one-time library initialisation outside the capture
            torch.manual_seed(2)
            try:
                gstep = GraphedTrainStep(rend, params, o, d, near, far, rgb, gt_feats=gt, background_rgb=bg,
                                         cos_anneal_ratio=1.0, warmup=2, grad_sync=sync, global_batch=B * world,
                                         ray_grads=pose)
                per_replay = int(gstep.launches_per_replay)
            except Exception as ex:                     # noqa: BLE001 - any capture failure means "run eagerly"
                print("CUDA graph capture failed, running eagerly: %r" % (ex,), file=sys.stderr)
                gstep = None
                torch.cuda.synchronize()

            def gfn(i):
                gstep(o, d, near, far, rgb, gt, bg)
        with ClockSampler(local) as cs:
            ms, launches = timed_steps(gfn if gstep else fn, args.steps, args.warmup, world, dev, flush, lib)
        if gstep:
            launches = per_replay * args.steps
        value = B * world * args.steps / (ms * 1e-3)
        # end to end through the public API with HOST buffers: pinned rays in, loss value out, every step
        host = [t.cpu().pin_memory() for t in (o, d, near, far, rgb)]
        dbuf = [torch.empty_like(t, device=dev) for t in host]
        barrier(world)
        t0 = time.perf_counter()
        n_e2e = max(2, args.steps // 2)
        # Every step copies its rays from pinned host memory and copies its loss back to pinned host memory; the host
        # READS the loss of step i after it has queued step i + 1 (one event per step), the way a training loop that logs
        # its loss without stalling the GPU does - all n_e2e values are read before the clock stops.
        loss_host = [torch.zeros(1).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        losses = []
        for i in range(n_e2e):
            for h, g_ in zip(host, dbuf):
                g_.copy_(h, non_blocking=True)
            torch.manual_seed(2 + i)
            if gstep:
                loss, _ = gstep(dbuf[0], dbuf[1], dbuf[2], dbuf[3], dbuf[4], gt, bg)
            else:
                loss, _ = train_step(rend, params, dbuf[0], dbuf[1], dbuf[2], dbuf[3], dbuf[4], gt_feats=gt,
                                     background_rgb=bg, cos_anneal_ratio=1.0, grad_sync=sync, global_batch=B * world)
            loss_host[i & 1].copy_(loss.reshape(1), non_blocking=True)      # device -> host copy of the step's result
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                losses.append(float(loss_host[(i - 1) & 1]))
        loss_ev[(n_e2e - 1) & 1].synchronize()
        losses.append(float(loss_host[(n_e2e - 1) & 1]))
        assert len(losses) == n_e2e and all(v == v for v in losses)
        barrier(world)
        e2e_t = max_over_ranks((time.perf_counter() - t0) / n_e2e, world, dev)
        # per-family device time of one extra step (CUDA events on the launching stream)
        import ctypes
        lib.vdn_prof_enable(1)
        fn(0)
        fam = {}
        for f, nm in ((0, "gemm_nt"), (1, "wgrad"), (2, "tc"), (3, "chain"), (4, "chain_train"), (5, "wgrad16")):
            msn, sp, fl, by = ctypes.c_double(), ctypes.c_longlong(), ctypes.c_double(), ctypes.c_double()
            lib.vdn_prof_read(f, ctypes.byref(msn), ctypes.byref(sp), ctypes.byref(fl))
            lib.vdn_prof_read_bytes(f, ctypes.byref(by))
            fam[nm] = (msn.value, sp.value, fl.value, by.value)
        lib.vdn_prof_enable(0)
        gemm_ms = sum(v[0] for v in fam.values())
        alg = alg_flops_per_ray(depth) * B      # (ray gradients of configs[4] add the small input-gradient GEMMs only)
        ach = alg / (gemm_ms * 1e-3) / 1e12
        tensor_view = {"achieved": ach, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_sustained"],
                       "algorithmic_flops_per_step": alg, "executed_flops_per_step": sum(v[2] for v in fam.values()),
                       "note": "algorithmic FLOPs of the whole step (SURVEY 8(d)) / summed MLP-kernel time",
                       "peak_source": pk["source"] + " bf16 sustained (kind::f16 rate)"}
        common = {"launches_by_family": {k: v[1] for k, v in fam.items()}, "ms_by_family": {k: v[0] for k, v in fam.items()},
                  "kernel_ms_per_step": gemm_ms, "share_of_step": gemm_ms / (ms / args.steps),
                  "all_families": {k: {"ms": v[0], "launches": v[1], "GB/s": (v[3] / (v[0] * 1e-3) / 1e9) if v[0] else 0.0,
                                       "TFLOP/s": (v[2] / (v[0] * 1e-3) / 1e12) if v[0] else 0.0}
                                   for k, v in fam.items()}}
        names = {"gemm_nt": "gemm_nt_kernel (FFMA)", "wgrad": "gemm_tn(_tc)_kernel (layer-wise weight gradient)",
                 "tc": "gemm_nt_tc_kernel (layer-wise tcgen05 kind::tf32 GEMM)",
                 "chain": "sdf_chain_tc_kernel (fused SDF value chain of the hierarchical sampler, kind::f16)",
                 "chain_train": "chain_kernel (fused training chains: SDF forward / normals / backward phases 1+2, colour and "
                                "NeRF forward / backward; tcgen05 kind::f16, A operand in TMEM, 16-bit saved tensors)",
                 "wgrad16": "wgrad16_kernel (grouped weight gradient, TMA tensor maps, MN-major fp16 operands)"}
        if args.precision == "tf32":
            dom = max(fam, key=lambda k: fam[k][0])
            d_ms, d_n, d_fl, d_by = fam[dom]
            gbs = d_by / (d_ms * 1e-3) / 1e9
            tfs = d_fl / (d_ms * 1e-3) / 1e12
            f_h, f_t = gbs / pk["hbm"], tfs / pk["bf16_sustained"]
            # the roof that binds the dominant kernel is the one it sits closer to
            if f_h >= f_t:
                roof = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": f_h,
                        "peak_source": pk["source"] + " hbm (copy)"}
            else:
                roof = {"bound": "tensor", "achieved": tfs, "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": f_t,
                        "peak_source": pk["source"] + " bf16 sustained"}
            roof.update({"kernel": "%s, %d launches per step" % (names[dom], d_n),
                         "hbm_frac": f_h, "tensor_frac": f_t,
                         # the committed capture is of the default workload (womsk_white, 512 rays): no figure for others
                         "traffic": ncu_traffic(dom) if (B == 512 and not depth and not pose) else None,
                         "traffic_note": "DRAM bytes per launch, mean of the launches of one step, from ncu --set full of "
                                         "this workload at 512 rays (profiles/r02_ncu_full_%s.txt); compare "
                                         "algorithmic_bytes_per_launch" % dom,
                         "algorithmic_bytes_per_launch": d_by / max(1, d_n), "algorithmic_bytes_per_step": d_by,
                         "executed_flops_per_launch": d_fl / max(1, d_n), "kernel_ms": d_ms,
                         "launch_ms": d_ms / max(1, d_n), "tensor_view": tensor_view})
        else:
            roof = dict(tensor_view)
            roof.update({"bound": "tensor", "traffic": None,
                         "kernel": "all MLP contractions of one step: gemm_nt_kernel (FFMA) x%d, gemm_tn_kernel (wgrad) x%d"
                         % (fam["gemm_nt"][1], fam["wgrad"][1])})
        roof.update(common)
        line.update({"metric": "train_rays_per_s", "unit": "rays/s", "value": value, "ms_per_step": ms / args.steps,
                     "config": {"workload": "womsk_white%s training step (BASELINE configs[%d]): render fwd + driver "
                                            "loss + bwd%s%s" % ("_wdepth" if depth else "", 2 if depth else (4 if pose else 1),
                                                                " incl. ray gradients (learnable poses)" if pose else "",
                                                                " + NCCL grad all-reduce" if world > 1 else ""),
                                "fused_chains": bool(ops.get_chain()) and args.precision == "tf32",
                                "operand_formats": "fp16 operands everywhere (cotangents carry a per-call power-of-two "
                                                   "loss scale), fp32 accumulate" if args.precision == "tf32" else "fp32",
                                "rays_per_step_per_gpu": B, "global_batch": B * world, "n_samples": 64,
                                "n_importance": 64, "n_outside": 32, "mode": args.precision, "cuda_graph": bool(gstep),
                                "l2": "256 MiB flush between timed iterations" + (" (+ one-element all-reduce to align the ranks, untimed)" if world > 1 else "")},
                     "e2e": {"value": B * world / e2e_t, "unit": "rays/s",
                             "h2d_bytes_per_step": int(sum(h.numel() * 4 for h in host)), "d2h_bytes_per_step": 4},
                     "roofline": roof})
        cpu_kind = "train"

    clocks = cs.summary()
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            from oracle import stage_ref
            have_ref = stage_ref.available()
            threads = os.cpu_count() or 1
            if cpu_kind == "grid":
                v, threads, times = cpu_reference_grid(1 << 17, 3)
                cb = {"value": v, "unit": "pts/s", "cores": threads, "kind": "port",
                      "sample": "131072 points x 3 reps of SDFNetwork.sdf, PyTorch CPU fp32 (oracle port)"}
            elif have_ref:
                from oracle import ref_runner
                r = ref_runner.run("cpu", B, 2, 1, depth, threads)
                cb = {"value": r["rays_per_s"], "unit": "rays/s", "cores": r["threads"], "kind": "reference",
                      "sample": "%d rays x 2 timed steps (1 warm-up) of render + loss + backward through the unmodified "
                                "reference classes (oracle/_ref), PyTorch CPU fp32 autograd" % B}
            else:
                v, threads, times = cpu_reference_step(128, depth, 2)
                cb = {"value": v, "unit": "rays/s", "cores": threads, "kind": "port",
                      "sample": "128 rays x 2 timed steps (1 warm-up) of render+loss+backward, PyTorch CPU fp32 "
                                "autograd (oracle port of the reference; oracle/_ref not staged on this box)"}
            line["cpu_baseline"] = cb
            if cpu_kind == "train" and not args.no_ref_cuda:
                # the reference's own eager CUDA path on this same GPU; reported, not a target
                line["reference_cuda_eager"] = reference_cuda_eager(B, depth)
        line.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
                     "scaling": "strong" if (cpu_kind == "grid" or args.global_batch > 0) else "weak", "vs_baseline": None,
                     "dtype": line.get("dtype", "f16" if args.precision == "tf32" else "f32"),
                     "data": "synthetic", "gpu_launches": int(launches), "clocks": clocks})
        order = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                 "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline",
                 "reference_cuda_eager", "clocks"]
        print(json.dumps({k: line[k] for k in order if k in line}))
    if world > 1:
        # A CUDA graph that captured NCCL collectives keeps communicator resources alive; tearing the process group down
        # under it was seen to block.  Results are printed: leave through a barrier and a hard exit instead.
        import torch.distributed as dist
        sys.stdout.flush()
        sys.stderr.flush()
        try:
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="train", choices=["train", "train_wdepth", "train_pose", "grid"],
                    help="train: BASELINE configs[1]; train_wdepth: configs[2] (use --global-batch 4096); grid: configs[3]; "
                         "train_pose: configs[4] (ray gradients; sweep --rays 2048..16384 per GPU)")
    ap.add_argument("--global-batch", type=int, default=0,
                    help="total rays per step, split evenly over the ranks (strong scaling); 0: --rays per GPU (weak)")
    ap.add_argument("--rays", type=int, default=512, help="rays per step per GPU")
    ap.add_argument("--resolution", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph-nccl", action="store_true",
                    help="multi-GPU: launch the step eagerly instead of replaying a CUDA graph that captured the step "
                         "together with its NCCL all-reduces")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference's eager-CUDA leg (subprocess)")
    ap.add_argument("--no-cuda-graph", action="store_true",
                    help="training workloads: launch every kernel eagerly instead of replaying the captured step")
    ap.add_argument("--precision", default="tf32", choices=["fp32", "tf32"],
                    help="fp32: exact FFMA kernels; tf32: tcgen05 tensor-core kernels (fp32 accumulate)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
