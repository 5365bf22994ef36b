#!/bin/bash
echo "== error vs chunk / mining (B=8192, D=512)"
for ch in 2048 1024; do timeout 600 python tests/diag_grad_error.py 512 8192 $ch usage,rand,hard,c3 2>&1 | grep "^B="; done
timeout 300 python tests/diag_grad_error.py 1024 8192 2048 usage,rand 2>&1 | grep "^B="
echo "== minings timing"; for m in usage rand relative grel; do python tests/tune_phases.py 8192 512 fp16x2 0 0 $m; done
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -12
echo "== bench"; timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 3000 gpurun_out/bench_r2_n1.json; tail -5 gpurun_out/bench_r2_n1.err
echo "== ncu selects"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_kernel --launch-skip 4 --launch-count 1 -o gpurun_out/ncu_lsel -f python tests/tune_phases.py 8192 512 fp16x2 0 0 relative > gpurun_out/ncu_lsel.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:select_kernel --launch-skip 12 --launch-count 3 -o gpurun_out/ncu_gsel -f python tests/tune_phases.py 8192 512 fp16x2 0 0 grel > gpurun_out/ncu_gsel.log 2>&1
