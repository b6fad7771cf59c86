# level-1 launch plans of the FUSED-policy step (every env outside PLAY phase is in MISC): "streams order" per line
run() { BGYM_L1_STREAMS_FUSED=$1 BGYM_L1_ORDER_FUSED=$2 timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print('[fused plan $1 $2] value %.3e kernel_ms %.4f fused %.3e graph-fused %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['graph_replay']['fused']))" || tail -3 gpurun_out/bench_v.err; }
while read a b; do run $a $b; done <<'PLANS'
0123456 2105463
0123456 3102456
0102100 3102456
0102200 3102456
0101100 3102456
0112000 3102456
PLANS
