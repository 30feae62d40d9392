#!/bin/bash
# scratch job of the current gpurun call (edited per call)
timeout 900 python -m pytest tests/test_gpu_sharding.py tests/test_gpu_large_golden.py -x -q 2>&1 | tail -3
AWM_TC=11x2 timeout 600 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -x -q 2>&1 | tail -3
for v in 11x2 12x2 11x2 12x2; do
AWM_TC=$v python bench.py --resident-only --steps 3 --warmup 2 > gpurun_out/r2m_bench_$v.json 2>gpurun_out/r2m_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/r2m_bench_$v.json"))
print("$v", d["ms_per_step"], d["payload_ok"], "stft", d["kernels"]["k_stft_mags_tc"]["ms_per_launch"])
PY
done
