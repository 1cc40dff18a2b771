N=$1
O=gpurun_out/mg; mkdir -p $O
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N "$@" > $O/${name}_${N}gpu.json 2> $O/${name}_${N}gpu.err; echo "$name N=$N exit $?"; }
run train --steps 30 --warmup 5 --no-cpu-baseline --no-ref-cuda
run train_pose_16384 --workload train_pose --rays 16384 --steps 3 --warmup 3 --no-cpu-baseline --no-ref-cuda
