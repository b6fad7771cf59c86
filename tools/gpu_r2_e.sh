#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "side_arrays or without_obs or empty_slab" 2>&1 | grep -v "^E    \+" | tail -15
PYTHONPATH=. python tools/exp/mirror_probe.py 2>&1 | tail -14
