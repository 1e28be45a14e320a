#!/bin/bash
# One GPU session: parity subset for the streaming kernel, bench line, ncu capture.
# usage: profiles/gpu_check.sh <tag> [pytest -k expression]
tag=${1:-x}
kexpr=${2:-dil}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$kexpr" 2>&1 | tail -6 | tee gpurun_out/${tag}_pytest.txt
timeout 600 python bench.py --steps 20 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 6 -c 2 -f -o gpurun_out/${tag}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
