#!/bin/bash
# Scratch job run on the GPU box through tools/gpurun_retry.sh; this default is the round-end check: per-phase timings, the GPU test suite, one bench line.
echo "== timings"; python tests/tune_phases.py 8192 512 fp16x2 0 0 usage
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
echo "== bench"; timeout 900 python bench.py | tail -c 2500
