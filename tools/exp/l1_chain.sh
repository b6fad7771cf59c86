# BGYM_L1_CHAIN modes on the production loop, same box
for m in 0 1 2 0 1 2; do
  BGYM_L1_CHAIN=$m timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print('[chain $m] value %.3e kernel_ms %.4f fused %.3e graph %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['graph_replay']['sampler_plus_step']))" || tail -3 gpurun_out/bench_v.err
done
