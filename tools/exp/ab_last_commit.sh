# A/B of the working tree against the last commit on the same box: the committed sources are built from a scratch copy
set -e
mkdir -p gpurun_out /tmp/ab
run() { timeout 300 python bench.py --steps 200 --warmup 20 --no-hands --no-cpu-baseline --no-ppo --no-facade --e2e-steps 3 > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); print('[$1] value %.3e kernel_ms %.4f fused %.3e graph %.3e' % (d['value'], d['roofline']['kernel_ms'], d['fused_rollout']['value'], d['graph_replay']['sampler_plus_step']))"; }
cp balatro_gym_b200/libbgym.so /tmp/ab/new.so
cp -r balatro_gym_b200/csrc /tmp/ab/csrc_new
cp -r tools/exp/_ab_old/* balatro_gym_b200/csrc/
python -c "import balatro_gym_b200 as b; b.build(force=True)"
run old; 
cp /tmp/ab/csrc_new/* balatro_gym_b200/csrc/; python -c "import balatro_gym_b200 as b; b.build(force=True)"
run new; 
cp -r tools/exp/_ab_old/* balatro_gym_b200/csrc/; python -c "import balatro_gym_b200 as b; b.build(force=True)"
run old
cp /tmp/ab/csrc_new/* balatro_gym_b200/csrc/; python -c "import balatro_gym_b200 as b; b.build(force=True)"
run new
