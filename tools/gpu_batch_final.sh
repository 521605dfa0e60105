#!/bin/bash
# Final-tree evidence in one gpurun call: GPU tests, smoke, both bench arms, the K1p capture behind roofline.traffic.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2_pytest_final.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_pytest_final.log; tail -3 $O/r2_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/run_kernels.py scramble3 --iters 3 > $O/r2_k1p_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:scramble_pairs -s 2 -c 1 -o $O/r2_k1p_prof_final python tools/run_kernels.py scramble3 --iters 3 > $O/r2_k1p_ncu.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r2_bench_n1.json 2> $O/r2_bench_n1.err; echo "bench rc=$?"; tail -3 $O/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_bench_ref_n1.json 2> $O/r2_bench_ref_n1.err; echo "ref rc=$?"
