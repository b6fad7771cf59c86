#!/bin/bash
# A/B: prebuilt library variants (balatro_gym_b200/libbgym_<tag>.so) -> smoke + short bench each
mkdir -p gpurun_out
for tag in "$@"; do
  cp balatro_gym_b200/libbgym_$tag.so balatro_gym_b200/libbgym.so; touch balatro_gym_b200/libbgym.so
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$tag.log 2>&1 || { echo "$tag SMOKE FAILED"; tail -3 gpurun_out/smoke_$tag.log; }
  python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$tag.json')); print('$tag value %.3e frac %.3f kernel_ms %.3f fused %.3e' % (d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['fused_rollout']['value']))" || tail -3 gpurun_out/bench_$tag.err
done
