#!/bin/bash
sed -n '/^cat > \/tmp\/pt.py/,/^PY$/p' tools/gpu_policy3.sh > /tmp/mk.sh; bash /tmp/mk.sh
python -c "from balatro_gym_b200 import _lib; _lib.build_policy(force=True)"
PYTHONPATH=. timeout 120 python /tmp/pt.py
PYTHONPATH=. BGYM_POLICY_CLOCK=1 timeout 120 python /tmp/pt.py 2>&1 | grep clocks | tail -1
