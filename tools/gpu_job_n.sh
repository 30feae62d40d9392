#!/bin/bash
# N-GPU bench the way the driver launches it (+ stage trace of one more step): bash tools/gpu_job_n.sh <N>
N=${1:-2}
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2k_bench_n$N.json 2> gpurun_out/r2k_bench_n$N.err; echo "rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2k_bench_n$N.json"))
print(d["n_gpus"], d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e_f32"]["ms_per_step"], d["payload_ok"], d["sharded_equals_single_gpu"], d["host_wall_ms_per_step"])
PY
AWM_TRACE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 1 --warmup 1 --resident-only > /dev/null 2> gpurun_out/r2k_trace_n$N.err
grep "trace" gpurun_out/r2k_trace_n$N.err | grep "rank 0" | head -10
