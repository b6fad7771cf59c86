#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_rules_classifier.py -x -q -m gpu -k "score_hands or known_answers or rules" 2>&1 | tail -3
python - <<'PY'
import torch, bench, json
dev = torch.device("cuda:0")
r = bench.bench_hands_jokers(torch, dev, 6540.5)
print("hands_with_jokers %.3e hands/s  %.3f ms" % (r["value"], r["ms_per_launch"]))
PY
