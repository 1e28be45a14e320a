#!/bin/bash
# Final round-2 GPU session on one B200: all -m gpu tests, smoke, both bench arms, ncu launch list of
# the bench command, ncu --set full summaries of the default kernels, the variant sweep for the record.
tag=${1:-r02f}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv,noheader > gpurun_out/${tag}_gpu.txt
(time timeout 900 python -m pytest tests -q -m gpu --maxfail=8 --tb=short --durations=8 --timeout 150) > gpurun_out/${tag}_pytest_full.txt 2>&1
tail -14 gpurun_out/${tag}_pytest_full.txt | cut -c1-200 | tee gpurun_out/${tag}_pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 420 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
j = json.load(open("gpurun_out/${tag}_bench_n1.json"))
print("ms", j["ms_per_step"], "value", j["value"], j["roofline"]["kernels"], j["roofline"]["whole_step"])
print("e2e", j["e2e"]["value"], json.dumps(j.get("e2e_variants"))[:600])
print("parity", j.get("parity_vs_reference"), "batch", j.get("batch"))
for c in j.get("configs", []):
    print(" ", c.get("config", "")[:46], c.get("ms_per_frame"), c.get("row_ms"), c.get("col_ms"), c.get("frac"), c.get("error"))
print("lancir", json.dumps(j.get("lancir"))[:500])
PY
tail -3 gpurun_out/${tag}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_n1.json 2> gpurun_out/${tag}_bench_ref.err
cut -c1-200 gpurun_out/${tag}_bench_reference_n1.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_bench.log 2>&1
tail -3 gpurun_out/${tag}_launches.csv | cut -c1-260
bash profiles/gpu_r02_ncu.sh ${tag} keep=default cfg=cfg5 cfg=cfg4 | tail -30
out=gpurun_out/${tag}_sweep.jsonl; : > $out
for cfg in cfg3 cfg4 cfg5; do
  for v in 0 1 2; do
    timeout 120 python profiles/pass_times.py --cfg $cfg --var-h $v --var-v $v >> $out 2>> ${out}.err
  done
done
for cfg in cfg2 rgb; do timeout 120 python profiles/pass_times.py --cfg $cfg >> $out 2>> ${out}.err; done
cut -c1-230 $out
