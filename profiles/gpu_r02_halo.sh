#!/bin/bash
# Exchange schedules of the sharded path side by side (gpurun --gpus N): hardware parity of the chosen
# AVIRB200_OPT_OVERLAP_HALO modes, then the headline bench per mode.
# usage: profiles/gpu_r02_halo.sh <tag> <N> "<modes>"
tag=$1; n=${2:-2}; modes=${3:-"1 3"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${tag}_gpus.txt
(time AVIR_NCCL_OVERLAPS=$(echo $modes | tr ' ' ',') timeout 400 python -m pytest tests/test_gpu_nccl.py -q -m gpu --tb=short --timeout 300 -s) > gpurun_out/${tag}_nccl_pytest.txt 2>&1
grep -E "mismatches=|passed|failed|skipped|Error|error" gpurun_out/${tag}_nccl_pytest.txt | cut -c1-200 | tail -40
for rep in 1 2; do
for m in $modes; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n + 10 * m + rep)) \
      bench.py --gpus $n --steps 40 --warmup 5 --no-extras --no-cpu-baseline --halo-mode $m > gpurun_out/${tag}_bench_n${n}_halo${m}_rep$rep.json 2> gpurun_out/${tag}_bench_n${n}_halo${m}_rep$rep.err
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/${tag}_bench_n${n}_halo${m}_rep$rep.json") if l.startswith("{")][-1])
    print("N=$n halo-mode $m rep $rep", "ms", round(j["ms_per_step"], 4), "Mpix/s", round(j["value"]), "sharded_parity", j.get("sharded_parity", {}).get("mismatches"))
except Exception as e:
    print("N=$n halo-mode $m: no line:", e)
PY
done
done
