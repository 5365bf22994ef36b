#!/bin/bash
echo "== diag"; python tests/diag_pair_ragged.py 640 200 bf16x3; python tests/diag_pair_ragged.py 640 200 fp16x2
echo "== timings"; python tests/tune_phases.py 8192 512 fp16x2 0 0 usage
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q -x --deselect "tests/test_gpu_parity.py::test_pair_kernels_match_single_cta_bitwise" 2>&1 | tail -4
