#!/bin/bash
# GPU session: all -m gpu tests, bench line (e2e through the pipelined host call), per-pass
# timings of the scheduling variants for cfg3 (both mirrors) and cfg2.
tag=${1:-sweep}
mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests -x -q -m gpu) 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err
out=gpurun_out/${tag}_variants.jsonl; : > $out
for cfg in cfg3 cfg3f4 cfg2; do
  for v in 0 1 2 3; do
    AVIRB200_STREAM_VARIANT_H=$v AVIRB200_STREAM_VARIANT_V=$v timeout 300 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
  done
  AVIRB200_DISABLE_STREAM=1 timeout 300 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
done
cut -c1-260 $out
