#!/bin/bash
# scratch job of the current gpurun call (edited per call)
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/r2e_bench.json 2>gpurun_out/r2e_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2e_bench.json"))
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["payload_ok"], d["host_wall_ms_per_step"])
print({k: v["ms_per_launch"] for k, v in d["kernels"].items()})
print(d["add_get"])
PY
for p in 3072 6144; do AWM_PIECE=$p python bench.py --no-cpu-baseline --steps 2 --warmup 2 > gpurun_out/r2e_piece$p.json 2>/dev/null; python -c "
import json;d=json.load(open('gpurun_out/r2e_piece$p.json'));print('piece',$p,d['e2e']['ms_per_step'],d['add_get']['add']['e2e_ms_per_step'])"; done
