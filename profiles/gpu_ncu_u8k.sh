#!/bin/bash
# ncu --set full of the 8K->4K u8 passes (integer source row pass, integer-output column pass)
mkdir -p gpurun_out
timeout 100 ncu --set full --clock-control none --import-source on -k regex:pass_kernel -c 2 -f -o gpurun_out/u8k_prof \
    python profiles/pass_times.py --cfg u8k --n 1 > gpurun_out/u8k_ncu.log 2>&1
tail -3 gpurun_out/u8k_ncu.log
