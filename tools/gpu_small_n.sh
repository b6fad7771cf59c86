#!/bin/bash
# small-slab kernel (BGYM_SMALL_N) vs the five-launch split step at small n; parity first
BGYM_SMALL_N=100000 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "native or golden or ragged or sampler" 2>&1 | tail -2
for sn in 0 10000000; do for e in 32 1024 4096 16384 65536; do BGYM_SMALL_N=$sn timeout 200 python bench.py --envs $e --steps 300 --warmup 30 --no-hands --no-cpu-baseline --no-ppo 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('small_n=$sn envs', $e, 'value %.3e kernel_ms %.4f fused-policy %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value']))"; done; done
BGYM_SMALL_N=100000 PYTHONPATH=. python tools/exp/facade_speed.py
