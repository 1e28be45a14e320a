#!/bin/bash
# Short 1-GPU validation session: all -m gpu tests, smoke, the bench line.   usage: profiles/gpu_r02_check.sh <tag>
tag=${1:-r02m}
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -q -m gpu --maxfail=8 --tb=short --durations=5 --timeout 150) > gpurun_out/${tag}_pytest_full.txt 2>&1
tail -12 gpurun_out/${tag}_pytest_full.txt | cut -c1-200 | tee gpurun_out/${tag}_pytest_gpu.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.txt
timeout 420 python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
j = json.load(open("gpurun_out/${tag}_bench_n1.json"))
print("ms", j["ms_per_step"], "value", j["value"], j["roofline"]["whole_step"])
print("e2e", j["e2e"]["value"], "parity", j.get("parity_vs_reference"))
for c in j.get("configs", []):
    print(" ", c.get("config", "")[:46], c.get("ms_per_frame"), c.get("row_ms"), c.get("col_ms"), c.get("error"))
print("lancir", j.get("lancir", {}).get("ms_per_frame"))
PY
tail -3 gpurun_out/${tag}_bench.err
