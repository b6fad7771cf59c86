#!/bin/bash
for impl in split fused; do for e in 4096 16384 65536 262144; do BGYM_STEP_IMPL=$impl timeout 200 python bench.py --envs $e --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$impl envs', $e, 'value %.3e kernel_ms %.4f fused-policy %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value']))"; done; done
