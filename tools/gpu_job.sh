#!/bin/bash
# scratch job of the current gpurun call (edited per call)
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/r2l_bench.json 2>gpurun_out/r2l_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2l_bench.json"))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e_f32"]["ms_per_step"], d["payload_ok"], d["host_wall_ms_per_step"])
print({k: v["ms_per_launch"] for k, v in d["kernels"].items()})
print(d["add_get"])
PY
