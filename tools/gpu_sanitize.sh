#!/bin/bash
# compute-sanitizer over the smoke run (2048 envs x 64 fused steps on both launch structures + 4096 hands, checked against the oracle)
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|error" gpurun_out/sanitize_$tool.log | head -8
done
# memcheck over the small-slab, sampler, host-handle, hands, rollout-kernel, side-array / host-mirror and fused-policy tests
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rollout.py -x -q -m gpu \
  -k "ragged or empty or sampler or known or host_buffer or replayed or featurize or gae or masked or side_arrays or fused_policy_forward_matches or first_layer" > gpurun_out/sanitize_memcheck_tests.log 2>&1
echo "== memcheck tests rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_memcheck_tests.log | tail -3
