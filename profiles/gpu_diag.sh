#!/bin/bash
# Which GPU tests are slow?  Bounded.
mkdir -p gpurun_out
(time timeout 100 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "float64 and product" --durations=8) > gpurun_out/diag_f64.txt 2>&1
tail -15 gpurun_out/diag_f64.txt | cut -c1-200
(time timeout 150 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "(defE or f4E or dilE) and product" --durations=12) > gpurun_out/diag_errd.txt 2>&1
tail -20 gpurun_out/diag_errd.txt | cut -c1-200
