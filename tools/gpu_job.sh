#!/bin/bash
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== HL timing"; python tests/tune_phases.py 8192 512 fp16x2 0 0 usage
for cfg in HL C2 C3 C4 C5; do
echo "== bench $cfg N=1"; timeout 900 python bench.py --config $cfg --steps 30 --warmup 5 > gpurun_out/bench_r2_${cfg}_n1.json 2> gpurun_out/bench_r2_${cfg}_n1.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_${cfg}_n1.json").read().strip().splitlines()[-1])
    print("$cfg", "ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"]), "roof", round(d["roofline"]["frac"],3), round(d["roofline"]["frac_of_issued_mma"],3), "row_frac", round(d["hbm_kernels"]["row_pass_frac"],3)); print("   phases", {k: round(v*1e3,1) for k,v in d["phase_ms"].items()})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_r2_${cfg}_n1.err").read()[-2000:])
PY
done
echo "== reference arm HL (full k=1 shape)"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 | cut -c1-700
