#!/bin/bash
# One gpurun call that measures and checks the opt-in kernels against the default path (round-2 opener):
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tests/run_experiments.sh > gpurun_out/experiments.log 2>&1; tail -60 gpurun_out/experiments.log'
# Every step runs under its own timeout: the stream-K kernel spins on flags, a bug there must not hang the box.
cd "$(dirname "$0")/.."
echo "== per-phase timings (HL: B=8192, D=512, fp16x2) =="
for v in "NPAIR_NONE=1" "NPAIR_LSE_TILES=1" "NPAIR_LSE_TILES=2" "NPAIR_GRAD_ONE_EX2=1" "NPAIR_GRAD_STREAMK=1" "NPAIR_GRAD_STREAMK=1 NPAIR_GRAD_ONE_EX2=1" \
         "NPAIR_LSE_TILES=1 NPAIR_GRAD_STREAMK=1 NPAIR_GRAD_ONE_EX2=1"; do
  echo "--- $v"
  env $v timeout 120 python tests/tune_phases.py || echo "FAILED/timeout: $v"
done
echo "== bf16 mode =="
for v in "NPAIR_NONE=1" "NPAIR_GRAD_STREAMK=1 NPAIR_GRAD_ONE_EX2=1"; do
  echo "--- $v"; env $v timeout 120 python tests/tune_phases.py 8192 512 bf16 || echo "FAILED/timeout: $v"
done
echo "== correctness against the default path =="
NPAIR_RUN_EXPERIMENTAL=1 timeout 1200 python -m pytest tests/test_gpu_experimental.py -q 2>&1 | tail -15
echo "== device-resident step: two calls vs npair_forward_backward =="
for f in "" "--fused-step"; do timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline $f 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['step_call'], 'ms/step', d['ms_per_step'], 'samples/s', d['value'])"; done
echo "== e2e: serial H2D vs prefetch (double-buffered bottoms) =="
timeout 400 python bench.py --steps 100 --warmup 10 --e2e-prefetch --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'], 'e2e_prefetch', d.get('e2e_prefetch'))"
# multi-GPU (separate gpurun --gpus 2 call):
#   NPAIR_P2P_RECORDS=1 timeout 600 python -m pytest tests/test_multi_gpu.py -q
#   for v in NPAIR_NONE=1 NPAIR_P2P_RECORDS=1; do env $v timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 \
#       --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 20 --no-cpu-baseline | cut -c1-300; done
