#!/bin/bash
# One gpurun call that produces the ncu evidence of a round: launch list of a bench run + --set full captures of every kernel on the default path and of the selects.
R=${1:-r02}
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches_$R.bench.json 2> gpurun_out/launches_$R.err
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 20 --launch-count 5 -o gpurun_out/prof_${R}_main -f python tests/tune_phases.py 8192 512 fp16x2 0 0 usage > gpurun_out/prof_${R}_main.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:local_select --launch-skip 4 --launch-count 1 -o gpurun_out/prof_${R}_lsel -f python tests/tune_phases.py 8192 512 fp16x2 0 0 relative > gpurun_out/prof_${R}_lsel.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:select_kernel --launch-skip 12 --launch-count 3 -o gpurun_out/prof_${R}_gsel -f python tests/tune_phases.py 8192 512 fp16x2 0 0 grel > gpurun_out/prof_${R}_gsel.log 2>&1
ls -la gpurun_out/prof_${R}_*.ncu-rep gpurun_out/launches_$R.csv
