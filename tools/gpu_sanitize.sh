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
# the warp-cooperative paths (Immolate's compaction, the consumable tile's re-convergence, the joker interpreter's warp-made
# rolls) need consumables in play: the c4 generator soak on both launch structures, and the hands-vs-oracle test
for tool in memcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/soak_parity.py --envs 2048 --steps 48 > gpurun_out/sanitize_${tool}_soak.log 2>&1
  echo "== $tool soak (multi-pass) rc=$?"; grep -E "ERROR SUMMARY|mismatches" gpurun_out/sanitize_${tool}_soak.log | cut -c1-160 | tail -2
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/soak_parity.py --envs 1024 --steps 48 --one-launch > gpurun_out/sanitize_${tool}_soak1.log 2>&1
  echo "== $tool soak (one launch) rc=$?"; grep -E "ERROR SUMMARY|mismatches" gpurun_out/sanitize_${tool}_soak1.log | cut -c1-160 | tail -2
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "score_hands_cuda_vs_oracle or score_hands_replayed" > gpurun_out/sanitize_${tool}_hands.log 2>&1
  echo "== $tool hands rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_${tool}_hands.log | tail -2
done
