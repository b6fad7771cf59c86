#!/bin/bash
# Standard GPU visit: parity tests, smoke, select-only experiment, bench.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python tools/exp_select_only.py
python bench.py --steps 200 --warmup 20 > gpurun_out/bench.json 2> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value %.3e frac %.3f kernel_ms %.3f fused %.3e e2e %.3e hands %.3e hfrac %.3f cpu %.3e (%s cores)' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['e2e']['value'], d['hands']['value'], d['hands']['roofline']['frac'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))" || tail -5 gpurun_out/bench.err
