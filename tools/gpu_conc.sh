#!/bin/bash
# concurrent-pass step: correctness + grid split sweep
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
run() { env "$@" python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$*: value %.3e frac %.3f kernel_ms %.3f fused %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value']))"; }
run BGYM_SERIAL_GATHER=1
run BGYM_GRID_MAIN=3 BGYM_GRID_PLAY=2 BGYM_GRID_OTHER=1 BGYM_GRID_DISCARD=1
run BGYM_GRID_MAIN=2 BGYM_GRID_PLAY=2 BGYM_GRID_OTHER=1 BGYM_GRID_DISCARD=1
run BGYM_GRID_MAIN=4 BGYM_GRID_PLAY=2 BGYM_GRID_OTHER=1 BGYM_GRID_DISCARD=1
run BGYM_GRID_MAIN=4 BGYM_GRID_PLAY=3 BGYM_GRID_OTHER=2 BGYM_GRID_DISCARD=1
run BGYM_GRID_MAIN=8 BGYM_GRID_PLAY=4 BGYM_GRID_OTHER=4 BGYM_GRID_DISCARD=4
run BGYM_GRID_MAIN=2 BGYM_GRID_PLAY=3 BGYM_GRID_OTHER=2 BGYM_GRID_DISCARD=1
