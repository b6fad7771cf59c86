# BGYM_LEVEL_KERNELS: 0 = one kernel per list (forked streams), 1 = level 2 as one kernel, 3 = both levels as one kernel each
for m in 0 1 3 0 1; do
  BGYM_LEVEL_KERNELS=$m timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print('[level kernels $m] value %.3e kernel_ms %.4f fused %.3e graph %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['graph_replay']['sampler_plus_step']))" || tail -3 gpurun_out/bench_v.err
done
BGYM_LEVEL_KERNELS=1 python tools/soak_parity.py --envs 16384 --steps 200 2>&1 | tail -1 | cut -c1-200
BGYM_LEVEL_KERNELS=3 python tools/soak_parity.py --envs 16384 --steps 200 2>&1 | tail -1 | cut -c1-200
