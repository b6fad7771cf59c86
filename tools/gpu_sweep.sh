#!/bin/bash
# Sweep the step-kernel variants: correctness (smoke vs oracle) + short bench each.
mkdir -p gpurun_out
for v in ${VARIANTS:-0 2 3 4 5}; do
  BGYM_VARIANT=$v python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v$v.log 2>&1 || { echo "variant $v SMOKE FAILED"; tail -5 gpurun_out/smoke_v$v.log; }
  BGYM_VARIANT=$v python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline > gpurun_out/bench_v$v.json 2> gpurun_out/bench_v$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v$v.json')); print('V$v value %.3e frac %.3f kernel_ms %.3f fused %.3e e2e %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['e2e']['value']))" || tail -3 gpurun_out/bench_v$v.err
done
