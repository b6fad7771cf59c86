#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_adapters.py tests/test_gpu_parity.py -x -q -m gpu -k "facade or sb3 or host_buffer or host_handle or golden_through or plain_c or adapters or unseeded" 2>&1 | tail -2
python - <<'PY'
import bench
print(bench.bench_host_facing())
PY
tools/gpu_prof_part.sh
