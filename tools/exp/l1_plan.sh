# level-1 launch plans on the production loop, same box: "streams order" per line (indexed / listing PLAY CONS GEN MISC DISCARD SHOP BLIND)
run() { BGYM_L1_STREAMS=$1 BGYM_L1_ORDER=$2 BGYM_L1_STREAMS_FUSED=$3 timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print('[$1 $2 | fused $3] value %.3e kernel_ms %.4f fused %.3e graph %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['graph_replay']['sampler_plus_step']))" || tail -3 gpurun_out/bench_v.err; }
while read a b c; do run $a $b $c; done <<'PLANS'
0121230 2105463 0123456
0121230 2150463 0123456
0121230 2510463 0123456
0121230 5210463 0123456
0122130 2105463 0123456
0121330 2105463 0123456
0121430 2105463 0123456
0121234 2105463 0123456
0121230 2105463 0123456
PLANS
