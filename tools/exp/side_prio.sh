# stream priorities of the forked level-1 streams on the production loop (BGYM_SIDE_PRIO: digit i = 1 -> forked stream i high priority)
for p in 000000 010000 110000 010000 000000; do
  BGYM_SIDE_PRIO=$p timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print('[side prio $p] value %.3e kernel_ms %.4f fused %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value']))" || tail -3 gpurun_out/bench_v.err
done
