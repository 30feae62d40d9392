#!/bin/bash
# scratch job of the current gpurun call (edited per call)
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_large_golden.py -x -q 2>&1 | tail -3
for v in pair single; do
AWM_VITERBI=$v python bench.py --resident-only --steps 3 --warmup 2 > gpurun_out/r2n_bench_$v.json 2>gpurun_out/r2n_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2n_bench_$v.json"))
print("$v", d["ms_per_step"], d["payload_ok"], "viterbi", d["kernels"]["k_viterbi"]["ms_per_launch"])
PY
done
