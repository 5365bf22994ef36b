#!/bin/bash
echo "== timings"; python tests/tune_phases.py 8192 512 fp16x2 0 0 usage; python tests/tune_phases.py 2048 512 fp16x2 0 0 usage; python tests/tune_phases.py 4096 512 bf16 0 0 usage
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
