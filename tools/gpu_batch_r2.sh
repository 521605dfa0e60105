#!/bin/bash
# One gpurun call of round 2: tests, timings, ncu captures (every ncu run directly behind the same command's plain run).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest7.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest7.log; tail -3 $O/r2_pytest7.log
timeout 300 python tools/depth_sweep.py 3 > $O/r2_depth_sweep_s3.log 2>&1; tail -6 $O/r2_depth_sweep_s3.log
python tools/run_kernels.py scramble3 --iters 3 > $O/r2_k1p_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:scramble_pairs -s 2 -c 1 -o $O/r2_k1p_prof python tools/run_kernels.py scramble3 --iters 3 > $O/r2_k1p_ncu.log 2>&1
cat $O/r2_k1p_plain.log
python tools/run_kernels.py small --iters 2 > $O/r2_small_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'prefix_kernel|adi_targets|decode2|decode3|scramble_sliced|scramble_pairs' -c 28 -o $O/r2_small_prof python tools/run_kernels.py small --iters 2 > $O/r2_small_ncu.log 2>&1
cat $O/r2_small_plain.log
python bench.py --steps 20 --warmup 5 --skip-other --skip-cpu > $O/r2_bench_short.json 2> $O/r2_bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_bench_n1.csv python bench.py --steps 20 --warmup 5 --skip-other --skip-cpu > $O/r2_bench_ncu.log 2>&1
echo "launch list rc=$?"
