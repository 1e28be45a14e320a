#!/bin/bash
# Round 2, first GPU session: scheduling variants of the streaming kernel (0 ring windows,
# 1 register windows, 2 register windows + TMA-staged column pass), both passes, four chains.
# Same out_sha1 across the variants of a config = same bits.
out=gpurun_out/r02_sweep1.jsonl; : > $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader > gpurun_out/r02_sweep1_gpu.txt
for cfg in cfg3 cfg3f4 cfg4 u8k u8kdil; do
  for v in 0 1 2; do
    AVIRB200_STREAM_VARIANT_H=$v AVIRB200_STREAM_VARIANT_V=$v timeout 300 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
  done
done
cut -c1-300 $out
tail -5 ${out}.err
