#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the NCCL / mailbox sharded path against the 1-GPU bits, then
# bench.py at 1 .. N GPUs.   usage: profiles/gpu_r02_multi.sh <tag> <N> ["<list of GPU counts to bench>"]
tag=$1; N=${2:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${tag}_gpus.txt
(time timeout 600 python -m pytest tests/test_gpu_nccl.py -q -m gpu --tb=short --timeout 300 -s) > gpurun_out/${tag}_nccl_pytest.txt 2>&1
grep -E "mismatches=|passed|failed|skipped|Error|error" gpurun_out/${tag}_nccl_pytest.txt | cut -c1-200 | tail -40
for n in ${3:-1 2 4 8}; do
  [ $n -le $N ] || continue
  if [ $n = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 20 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
        bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  fi
  python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/${tag}_bench_n$n.json") if l.startswith("{")][-1])
    print("N=$n", "ms", round(j["ms_per_step"], 4), "Mpix/s", round(j["value"]), "e2e", round(j["e2e"]["value"]), "sharded_parity", j.get("sharded_parity"))
    for k, v in (j.get("multi_gpu_configs") or {}).items():
        print("   ", k, json.dumps(v)[:300])
except Exception as e:
    print("N=$n: no line:", e)
PY
  tail -2 gpurun_out/${tag}_bench_n$n.err | cut -c1-300
  # the exchange schedules side by side (headline only)
  if [ $n -gt 1 ] && [ -n "$HALO_MODES" ]; then
    for m in $HALO_MODES; do
      timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n + 10 * m)) \
          bench.py --gpus $n --steps 40 --warmup 5 --no-extras --no-cpu-baseline --halo-mode $m > gpurun_out/${tag}_bench_n${n}_halo$m.json 2> gpurun_out/${tag}_bench_n${n}_halo$m.err
      python - <<PY
import json
try:
    j = json.loads([l for l in open("gpurun_out/${tag}_bench_n${n}_halo$m.json") if l.startswith("{")][-1])
    print("N=$n halo-mode $m", "ms", round(j["ms_per_step"], 4), "Mpix/s", round(j["value"]), "sharded_parity", j.get("sharded_parity"))
except Exception as e:
    print("N=$n halo-mode $m: no line:", e)
PY
    done
  fi
done
