#!/bin/bash
# One GPU-box session: tcgen05 probe, the -m gpu parity suite (one pytest process per group so that a faulting
# kernel cannot poison the CUDA context of the others), smoke(), bench lines and an ncu launch list.
# Usage (from the repo root, under gpurun):  bash tools/gpu_ci.sh [quick]
set -u
mkdir -p gpurun_out
OUT=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/gpu.txt 2>&1
echo "== probe_tc" ; timeout 120 ./build/probe_tc > $OUT/probe.log 2>&1; echo "probe exit $?"; tail -14 $OUT/probe.log

TESTGROUPS=("(embedder or packed) and not tf32" "sdf_forward and not tf32" "(backward_vs_fp64 or depth_head) and not tf32"
        "nerf and not tf32" "(upsample or cat_z) and not tf32" "(render_core or composite) and not tf32"
        "(full_render or without_background) and not tf32" "extract_fields and not tf32" "full_size and not tf32" "tf32")
: > $OUT/pytest.log
i=0
for g in "${TESTGROUPS[@]}"; do
  i=$((i+1))
  echo "== pytest group $i: $g" | tee -a $OUT/pytest.log
  timeout 900 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider -k "$g" >> $OUT/pytest.log 2>&1
  echo "group $i exit $?" | tee -a $OUT/pytest.log
  tail -3 $OUT/pytest.log
done
grep -E "passed|failed|error" $OUT/pytest.log | tail -12

echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -3 $OUT/smoke.log
if [ "${1:-}" != "quick" ]; then
  echo "== bench train tf32"; timeout 900 python bench.py --precision tf32 --steps 10 --warmup 3 > $OUT/bench_train_tf32.json 2> $OUT/bench_train_tf32.err; echo "exit $?"; tail -c 1800 $OUT/bench_train_tf32.json; tail -3 $OUT/bench_train_tf32.err
  echo "== bench grid tf32"; timeout 900 python bench.py --precision tf32 --workload grid --steps 3 --warmup 3 > $OUT/bench_grid_tf32.json 2> $OUT/bench_grid_tf32.err; echo "exit $?"; tail -c 1500 $OUT/bench_grid_tf32.json; tail -3 $OUT/bench_grid_tf32.err
  echo "== bench train"; timeout 900 python bench.py --precision fp32 --steps 10 --warmup 3 > $OUT/bench_train.json 2> $OUT/bench_train.err; echo "exit $?"; tail -c 1800 $OUT/bench_train.json; tail -3 $OUT/bench_train.err
  echo "== bench grid"; timeout 900 python bench.py --precision fp32 --workload grid --steps 2 --warmup 3 > $OUT/bench_grid.json 2> $OUT/bench_grid.err; echo "exit $?"; tail -c 1500 $OUT/bench_grid.json; tail -3 $OUT/bench_grid.err
  echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; echo "exit $?"; tail -c 900 $OUT/bench_ref.json
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_train.csv \
      python bench.py --precision tf32 --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > $OUT/ncu_train.log 2>&1; echo "ncu exit $?"
  wc -l $OUT/launches_train.csv
  echo "== ncu full captures (one launch of each tcgen05 kernel)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:sdf_chain_tc -c 1 -f -o $OUT/prof_chain \
      python bench.py --precision tf32 --workload grid --resolution 256 --steps 1 --warmup 3 --no-cpu-baseline > $OUT/ncu_chain.log 2>&1; echo "ncu chain exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc -s 100 -c 2 -f -o $OUT/prof_nt \
      python bench.py --precision tf32 --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > $OUT/ncu_nt.log 2>&1; echo "ncu nt exit $?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_tc -s 40 -c 2 -f -o $OUT/prof_tn \
      python bench.py --precision tf32 --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graph > $OUT/ncu_tn.log 2>&1; echo "ncu tn exit $?"
fi
echo "== done"
