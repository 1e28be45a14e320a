#!/bin/bash
# ncu --set full of one row-pass and one column-pass launch per argument; the summaries
# (profiles/ncu_summary.py: raw metrics, opcode mix, stall reasons, hottest SASS) come back as text,
# the reports themselves only for arguments marked keep= (gpurun_out is capped at 64 MiB).
# usage: profiles/gpu_r02_ncu.sh <tag> 0:0 keep=4:4 cfg=u8k ...   ("h:v" = scheduling variants of cfg3)
tag=$1; shift
mkdir -p gpurun_out
for a in "$@"; do
  keep=0
  case $a in keep=*) keep=1; a=${a#keep=};; esac
  case $a in
    cfg=*) cfg=${a#cfg=}; vh=-1; vv=-1; name=${tag}_${cfg};;
    default) cfg=cfg3; vh=-1; vv=-1; name=${tag}_cfg3_default;;
    *) cfg=cfg3; vh=${a%%:*}; vv=${a##*:}; name=${tag}_cfg3_v${vh}_${vv};;
  esac
  # --only col: the script runs the row pass once (the column pass's input), then column passes
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:pass_kernel -c 2 -f -o gpurun_out/$name \
      python profiles/pass_times.py --cfg $cfg --n 1 --only col --var-h $vh --var-v $vv > gpurun_out/${name}.log 2>&1
  python profiles/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_ncu_summary.txt 2>&1
  [ $keep = 1 ] || rm -f gpurun_out/$name.ncu-rep
  grep -E "Kernel Name|gpu__time_duration|pipe_fma_cycles|issue_active|stalls:" gpurun_out/${name}_ncu_summary.txt | cut -c1-220
done
ls -la gpurun_out/ | awk '{print $5, $9}'
