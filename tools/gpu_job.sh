#!/bin/bash
echo "== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches_r02.bench.json 2> gpurun_out/launches_r02.err; tail -2 gpurun_out/launches_r02.err
for cfg in HL C5 C3 C4 C2; do
extra=""; if [ $cfg = C4 ] || [ $cfg = C2 ]; then extra="--no-cpu-baseline"; fi
echo "== bench $cfg N=1"; timeout 900 python bench.py --config $cfg $extra > gpurun_out/bench_r2_${cfg}_n1.json 2> gpurun_out/bench_r2_${cfg}_n1.err; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_${cfg}_n1.json").read().strip().splitlines()[-1])
    print("$cfg ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "other", round(d["step_call_other"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "cpu", d["cpu_baseline"] and round(d["cpu_baseline"]["value"],1), "roof", d["roofline"]["frac"], "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    print("   phases", {k: round(v*1e3,1) for k,v in d["phase_ms"].items() if v>0})
    print("   other_minings", d.get("other_minings") and {k:(round(v["ms_per_step"],4), round(v["thresholds_select_ms"]*1e3,1)) for k,v in d["other_minings"].items()})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_r2_${cfg}_n1.err").read()[-1500:])
PY
done
