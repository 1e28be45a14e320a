#!/bin/bash
# Times scheduling variants of the streaming kernel on one config (one process each; the row
# pass takes the variant of AVIRB200_STREAM_VARIANT_H, the column pass of ..._V).
# usage: profiles/variant_sweep.sh <out.jsonl> <cfg> "<row variants>" "<column variants>"
out=$1; cfg=${2:-cfg3}; hv=${3:-0}; vv=${4:-8}
: > $out
for v in $hv; do
  AVIRB200_STREAM_VARIANT_H=$v AVIRB200_STREAM_VARIANT_V=8 timeout 300 python profiles/pass_times.py --cfg $cfg --only row >> $out 2>> ${out}.err
done
for v in $vv; do
  AVIRB200_STREAM_VARIANT_H=0 AVIRB200_STREAM_VARIANT_V=$v timeout 300 python profiles/pass_times.py --cfg $cfg --only col >> $out 2>> ${out}.err
done
AVIRB200_DISABLE_STREAM=1 timeout 300 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err
cut -c1-330 $out
