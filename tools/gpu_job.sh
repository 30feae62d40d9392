#!/bin/bash
# scratch job of the current gpurun call (edited per call)
timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_sharding.py -x -q 2>&1 | tail -3
AWM_TRACE=1 python tools/trace_add.py 12288 6144 4096 3072 2>&1 | grep -E "embed pipeline|add_s16|copy" | awk '/embed pipeline/ {last=$0} /add_s16/ {print last; print} /copy/ {print}'
