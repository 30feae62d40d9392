#!/bin/bash
# One gpurun call: GPU test suite, bench line (both arms), ncu launch list of the resident legs of the bench command, ncu --set full of
# one 10 min add + get.   usage (from the repo root, on the GPU box): bash tools/gpu_check.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_gputests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_gputests.log
tail -4 gpurun_out/${tag}_gputests.log
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --resident-only > gpurun_out/${tag}_launches_bench.log 2>&1; echo "ncu rc=$?"
bash tools/gpu_ncu_full.sh ${tag} 24
head -c 400 gpurun_out/${tag}_bench.json
