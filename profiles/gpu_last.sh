#!/bin/bash
# Last session of the round: all -m gpu tests, smoke, bench line, per-config timings.
tag=${1:-last}
mkdir -p gpurun_out
(time timeout 500 python -m pytest tests -q -m gpu --maxfail=5 --tb=short --durations=5) > gpurun_out/${tag}_pytest_full.txt 2>&1
tail -22 gpurun_out/${tag}_pytest_full.txt | cut -c1-250 | tee gpurun_out/${tag}_pytest.txt
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-300 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 200 python profiles/bench_configs.py > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
cut -c1-200 gpurun_out/${tag}_configs.jsonl
timeout 100 python profiles/pass_times.py --cfg cfg2 > gpurun_out/${tag}_passes.jsonl 2> gpurun_out/${tag}_passes.err; cut -c1-260 gpurun_out/${tag}_passes.jsonl
