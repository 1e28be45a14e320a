#!/bin/bash
# Full round-end style GPU session: all -m gpu tests, smoke, both bench arms, ncu launch list,
# per-config and per-pass timings.
# usage: profiles/gpu_full.sh <tag>
tag=${1:-full}
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -q -m gpu --maxfail=5 --tb=short --durations=8) > gpurun_out/${tag}_pytest_full.txt 2>&1
tail -25 gpurun_out/${tag}_pytest_full.txt | cut -c1-250 | tee gpurun_out/${tag}_pytest.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
cut -c1-300 gpurun_out/${tag}_bench_ref.json; tail -3 gpurun_out/${tag}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -3 gpurun_out/${tag}_launches.csv | cut -c1-300
timeout 300 python profiles/bench_configs.py > gpurun_out/${tag}_configs.jsonl 2> gpurun_out/${tag}_configs.err
cut -c1-200 gpurun_out/${tag}_configs.jsonl
out=gpurun_out/${tag}_passes.jsonl; : > $out
for cfg in u8k cfg4; do timeout 200 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err; done
cut -c1-260 $out
