#!/bin/bash
# scratch job of the current gpurun call (edited per call)
timeout 900 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_sharding.py tests/test_gpu_zz_fullsize.py -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline > gpurun_out/r2j_bench.json 2>gpurun_out/r2j_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2j_bench.json"))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_f32"]["ms_per_step"], d["payload_ok"], d["host_wall_ms_per_step"])
print(d["add_get"])
PY
