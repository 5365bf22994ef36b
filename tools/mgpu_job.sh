#!/bin/bash
# usage: bash tools/mgpu_job.sh N "HL C4 C5" [run_tests]
N=$1; CFGS=$2; TESTS=$3
if [ -n "$TESTS" ]; then echo "== multi-gpu tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -s 2>&1 | grep -E "mgpu|passed|failed|Error|error" | head -40; fi
port=29700
for cfg in $CFGS; do
port=$((port+1))
echo "== bench $cfg N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --config $cfg --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_${cfg}_n$N.json 2> gpurun_out/bench_r2_${cfg}_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_r2_${cfg}_n$N.json").read().strip().splitlines()[-1])
    print("$cfg N=$N ms/step", round(d["ms_per_step"],4), "value", round(d["value"]), "e2e", round(d["e2e"]["value"])); print("   phases", {k: round(v*1e3,1) for k,v in d["phase_ms"].items()}); print("   variants", d.get("sharded_variants")); print("   parity", d["parity_check"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/bench_r2_${cfg}_n$N.err").read()[-3000:])
PY
done
