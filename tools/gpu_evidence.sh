#!/bin/bash
# Round-2 evidence session on one B200 (run under gpurun from the repo root): parity suite, smoke, bench lines of every
# driver-visible workload, the ncu launch list and `--set full` captures, all condensed ON THE BOX into small text files
# under gpurun_out/ev/ (the .ncu-rep files are deleted: gpurun_out/ is limited to 64 MiB).  profiles/ gets copies.
# EV_NCU_ONLY=1 skips the tests and bench lines and only redoes the ncu captures.
set -u
O=gpurun_out/ev
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $O/gpu.txt 2>&1
if [ "${EV_NCU_ONLY:-0}" != "1" ]; then
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q > $O/pytest_gpu.log 2>&1; tail -2 $O/pytest_gpu.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?"; tail -2 $O/smoke.log
b() { name=$1; shift; timeout 1200 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err; echo "bench $name exit $?"; tail -c 400 $O/bench_$name.json; echo; }
b train --steps 20 --warmup 5
b reference_cpu --impl reference --steps 2 --warmup 1
b train_fp32 --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda
b train_eager --no-cuda-graph --steps 10 --warmup 3 --no-cpu-baseline --no-ref-cuda
b train_wdepth_512 --workload train_wdepth --steps 10 --warmup 3 --no-cpu-baseline
b train_wdepth_4096 --workload train_wdepth --rays 4096 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda
b train_pose_2048 --workload train_pose --rays 2048 --steps 5 --warmup 3 --no-cpu-baseline --no-ref-cuda
b grid --workload grid --steps 3 --warmup 3 --no-cpu-baseline
fi
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-ref-cuda"
echo "== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda > $O/ncu_launches.log 2>&1
python tools/summarize_launches.py $O/launches_train.csv > $O/ncu_launches_train.txt 2>&1; head -12 $O/ncu_launches_train.txt
full() {  # name, kernel regex, skip, count, extra bench args
  name=$1; rx=$2; skip=$3; cnt=$4; shift 4
  timeout 1500 ncu --set full --import-source on --clock-control none -k "regex:$rx" --launch-skip $skip --launch-count $cnt -o $O/full_$name -f $BENCH "$@" > $O/ncu_full_$name.log 2>&1
  ncu -i $O/full_$name.ncu-rep --page raw --csv 2>/dev/null | python tools/summarize_ncu.py > $O/ncu_full_$name.txt
  ncu -i $O/full_$name.ncu-rep --page source --csv --print-source sass 2>/dev/null | python tools/summarize_ncu_source.py > $O/ncu_source_$name.txt
  rm -f $O/full_$name.ncu-rep
  grep -c "== launch" $O/ncu_full_$name.txt
}
echo "== ncu full"
full chain_train chain_kernel 8 8
full wgrad16 wgrad16_kernel 3 3
full chain sdf_chain_tc_kernel 5 5
full pointwise_512 'composite_fwd_kernel|composite_bwd_kernel|upsample_kernel' 6 6
full pointwise_16k 'composite_fwd_kernel|composite_bwd_kernel|upsample_kernel' 6 6 --rays 16384
ls -la $O
