#!/bin/bash
# Quick GPU visit: rollout parity test + smoke + a short step bench (no hands / cpu legs).
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "native or golden or invariants" 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for i in 1 2; do
python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; python -c "
import json; d=json.load(open('gpurun_out/bench_q.json')); print('value %.3e frac %.3f kernel_ms %.3f fused %.3e e2e %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['e2e']['value']))" || tail -5 gpurun_out/bench_q.err
done
