#!/bin/bash
# Multi-GPU bench lines on ONE box: bash tools/gpu_multi.sh N   (run under `gpurun --gpus N`)
set -u
N=${1:-2}
O=gpurun_out/mg; mkdir -p $O
run() { name=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > $O/${name}_${N}gpu.json 2> $O/${name}_${N}gpu.err; echo "$name N=$N exit $?"; tail -c 600 $O/${name}_${N}gpu.json | head -c 600; echo; }
run train --steps 20 --warmup 5 --no-cpu-baseline --no-ref-cuda
run grid --workload grid --steps 3 --warmup 3 --no-cpu-baseline
run train_wdepth_gb4096 --workload train_wdepth --global-batch 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda
run train_pose_2048 --workload train_pose --rays 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda
