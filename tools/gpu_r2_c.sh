#!/bin/bash
# full GPU suite + smoke + bench (both arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value %.3e frac %.3f kernel_ms %.3f fused %.3e graph %.3e e2e %.3e hands %.3e hj %.3e ppo %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['graph_replay']['sampler_plus_step'], d['e2e']['value'], d['hands']['value'], d['hands']['with_jokers']['value'], d['ppo_rollout']['value']))" || tail -5 gpurun_out/bench.err
