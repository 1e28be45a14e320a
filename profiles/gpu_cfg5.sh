#!/bin/bash
# GPU session: per-pass timings of cfg5 / 8K->4K u8 with the streaming kernel on and off,
# all -m gpu tests with durations, bench line.
tag=${1:-cfg5}
mkdir -p gpurun_out
out=gpurun_out/${tag}_passes.jsonl; : > $out
for cfg in cfg5 u8k u8kdil cfg4; do
  timeout 200 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
  AVIRB200_DISABLE_STREAM=1 timeout 200 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
done
cut -c1-260 $out; tail -3 ${out}.err
(time timeout 600 python -m pytest tests -q -m gpu --maxfail=5 --tb=short --durations=12) > gpurun_out/${tag}_pytest_full.txt 2>&1; tail -40 gpurun_out/${tag}_pytest_full.txt | cut -c1-250
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-400 gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
