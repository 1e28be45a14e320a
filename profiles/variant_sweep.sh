#!/bin/bash
# Times every scheduling variant of the streaming kernel on one config (one process each).
# usage: profiles/variant_sweep.sh <out.jsonl> [cfg] [variants...]
out=$1; cfg=${2:-cfg3}; shift; shift
vars=${@:-0 1 2 3 4 5 6 7}
: > $out
for v in $vars; do
  AVIRB200_STREAM_VARIANT=$v timeout 300 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
done
AVIRB200_DISABLE_STREAM=1 timeout 300 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
cat $out
