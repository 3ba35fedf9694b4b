#!/bin/bash
# A/B harness on the GPU box: for every library variant given, report decoder throughput at 2^22 / 2^16 and ICP launch times.
for lib in "$@"; do
  echo "=== $lib"
  DIF_LIB_PATH=$lib timeout 200 python tools/tc_timing.py 22 2>&1 | head -3
  DIF_LIB_PATH=$lib timeout 200 python tools/icp_time.py 2>&1 | head -2
  DIF_LIB_PATH=$lib timeout 300 python bench.py --no-sweep --cpu-sample 1 --steps 100 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('bench', round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), d['roofline']['other'])"
done
