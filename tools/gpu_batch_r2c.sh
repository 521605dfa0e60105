#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest9.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest9.log; tail -3 $O/r2_pytest9.log
timeout 300 python tools/depth_sweep.py > $O/r2_depth_sweep_c.log 2>&1; grep -E "depth (320|321|480|1000)" $O/r2_depth_sweep_c.log
CUBE_SLICE=240 timeout 120 python tools/run_kernels.py scramble3_d1000 2>&1 | head -2
CUBE_SLICE=176 timeout 120 python tools/run_kernels.py scramble3_d1000 2>&1 | head -2
CUBE_SLICE=80 timeout 120 python tools/run_kernels.py scramble3_d1000 2>&1 | head -2
python tools/run_kernels.py small --iters 2 > $O/r2_small_plain_c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:'decode2|scramble_sliced' -c 8 -o $O/r2_small_prof_c python tools/run_kernels.py small --iters 2 > $O/r2_small_ncu_c.log 2>&1
cat $O/r2_small_plain_c.log
python tools/run_kernels.py scramble3 --iters 3 > $O/r2_k1p_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:scramble_pairs -s 2 -c 1 -o $O/r2_k1p_prof_c python tools/run_kernels.py scramble3 --iters 3 > $O/r2_k1p_ncu.log 2>&1
