#!/bin/bash
# training step at larger per-GPU batches than BASELINE's 8 crops: how much of the 8-crop step is launch / latency bound
for b in 8 16 32 64; do
  timeout 300 python bench.py --workload train --batch $b --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 2 2>/dev/null | \
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('batch $b:', round(d['value'],1), 'MPix/s', round(d['ms_per_step'],1), 'ms per step', round(d['roofline']['achieved'],1), 'TFLOP/s algorithmic')"
done
