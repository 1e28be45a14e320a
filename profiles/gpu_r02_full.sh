#!/bin/bash
# Round 2 GPU session: all -m gpu tests, smoke, scheduling-variant sweep (same out_sha1 across the
# variants of a config = same bits), both bench arms, ncu launch list of the bench command.
# usage: profiles/gpu_r02_full.sh <tag> [skip-tests]
tag=${1:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv,noheader > gpurun_out/${tag}_gpu.txt
if [ -z "$2" ]; then
  (time timeout 900 python -m pytest tests -q -m gpu --maxfail=8 --tb=short --durations=8 --timeout 150) > gpurun_out/${tag}_pytest_full.txt 2>&1
  tail -30 gpurun_out/${tag}_pytest_full.txt | cut -c1-250 | tee gpurun_out/${tag}_pytest.txt
  timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
fi
out=gpurun_out/${tag}_sweep.jsonl; : > $out
for cfg in cfg3 cfg3f4 cfg4 u8k u8kdil; do
  for v in 0 1 2; do
    timeout 120 python profiles/pass_times.py --cfg $cfg --var-h $v --var-v $v >> $out 2>> ${out}.err
  done
done
for v in 0 1 2; do
  timeout 120 python profiles/pass_times.py --cfg cfg5 --var-h $v --var-v $v >> $out 2>> ${out}.err
done
for cfg in cfg2 rgb; do
  timeout 120 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
done
timeout 120 python profiles/pass_times.py --cfg cfg5 --family 2 >> $out 2>> ${out}.err   # tile kernel, for reference
timeout 120 python profiles/pass_times.py --cfg cfg2 --family 2 >> $out 2>> ${out}.err
timeout 120 python profiles/pass_times.py --cfg cfg2 --all-chains 1 >> $out 2>> ${out}.err
cut -c1-260 $out; tail -3 ${out}.err
timeout 420 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cut -c1-6000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err
cut -c1-300 gpurun_out/${tag}_bench_ref.json; tail -3 gpurun_out/${tag}_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -4 gpurun_out/${tag}_launches.csv | cut -c1-300
