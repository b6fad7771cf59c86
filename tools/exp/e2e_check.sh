python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "side_arrays or mirror" 2>&1 | tail -2
for i in 1 2; do python bench.py --steps 100 --warmup 10 --no-hands --no-cpu-baseline --no-ppo --no-facade > gpurun_out/bench_v.json 2> gpurun_out/bench_v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_v.json')); e=d['e2e']; print('value %.3e e2e %.3e d2h %.1f MB pcie %.1f GB/s' % (d['value'], e['value'], e['d2h_bytes_per_step']/1e6, e['pcie_gbs_rank0']))" || tail -5 gpurun_out/bench_v.err; done
