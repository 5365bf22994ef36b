#!/bin/bash
echo "== timings"; python tests/tune_phases.py 8192 512 fp16x2 0 0 usage; python tests/tune_phases.py 8192 512 fp16x2 0 0 relative; NPAIR_LIB=npairloss_b200/lib/variant_minb4.so python tests/tune_phases.py 8192 512 fp16x2 0 0 relative
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
