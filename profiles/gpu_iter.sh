#!/bin/bash
# One GPU iteration: optional microbenchmarks, all -m gpu tests, bench line, per-config timings,
# one ncu --set full capture of the two pass kernels.
# usage: profiles/gpu_iter.sh <tag> [microbench binaries...]
tag=${1:-x}; shift
mkdir -p gpurun_out
for mb in "$@"; do timeout 120 $mb > gpurun_out/${tag}_$(basename $mb).jsonl 2>&1; cat gpurun_out/${tag}_$(basename $mb).jsonl; done
(time timeout 1500 python -m pytest tests -q -m gpu --maxfail=5 --tb=short) > gpurun_out/${tag}_pytest_full.txt 2>&1; tail -30 gpurun_out/${tag}_pytest_full.txt | cut -c1-300 | tee gpurun_out/${tag}_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 600 python profiles/bench_configs.py > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
cut -c1-200 gpurun_out/${tag}_configs.jsonl; tail -2 gpurun_out/${tag}_configs.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pass_kernel -s 6 -c 2 -f -o gpurun_out/${tag}_prof \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
