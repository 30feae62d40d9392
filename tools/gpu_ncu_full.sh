#!/bin/bash
# ncu --set full of the kernels of one 10 min add + get: bash tools/gpu_ncu_full.sh <tag> [launch count] [kernel regex]
# (the report is summarised on the box; it only travels back when it fits gpurun's 64 MiB limit)
tag=${1:-r2}
cnt=${2:-24}
re=${3:-"k_"}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$re" -c $cnt -o gpurun_out/${tag}_full -f python tools/ncu_target.py > gpurun_out/${tag}_full.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/${tag}_full.ncu-rep > gpurun_out/${tag}_full_summary.md
ls -la gpurun_out/${tag}_full.ncu-rep
sz=$(stat -c %s gpurun_out/${tag}_full.ncu-rep)
if [ "$sz" -gt 50000000 ]; then rm gpurun_out/${tag}_full.ncu-rep; echo "report dropped (too large), summary kept"; fi
