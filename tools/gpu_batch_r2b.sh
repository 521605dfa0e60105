#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest8.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest8.log; tail -3 $O/r2_pytest8.log
timeout 300 python tools/depth_sweep.py 3 > $O/r2_depth_sweep_s3b.log 2>&1; tail -5 $O/r2_depth_sweep_s3b.log
timeout 300 python tools/depth_sweep.py 2 > $O/r2_depth_sweep_s2b.log 2>&1; tail -4 $O/r2_depth_sweep_s2b.log
python tools/run_kernels.py scramble3 --iters 3 > $O/r2_k1p_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:scramble_pairs -s 2 -c 1 -o $O/r2_k1p_prof_b python tools/run_kernels.py scramble3 --iters 3 > $O/r2_k1p_ncu.log 2>&1
python tools/run_kernels.py small --iters 2 > $O/r2_small_plain_b.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'decode2|scramble_sliced' -c 8 -o $O/r2_small_prof_b python tools/run_kernels.py small --iters 2 > $O/r2_small_ncu_b.log 2>&1
cat $O/r2_small_plain_b.log
