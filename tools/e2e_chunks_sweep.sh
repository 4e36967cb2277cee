#!/bin/bash
# e2e of the default workload for several pipeline depths, with all host cores and confined to 4 cores (the share of one
# rank on an 8-GPU box with 32 hardware threads)
p() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'MPix/s device', round(d['e2e']['value'],1), 'e2e', round(d['e2e']['s_per_step']*1e3,1), 'ms')"; }
for c in 2 3 4; do CRDR_PIPELINE_CHUNKS=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 6 2>/dev/null | p "all cores, chunks $c:"; done
for c in 2 3 4 6; do CRDR_CODER_THREADS=4 CRDR_PIPELINE_CHUNKS=$c taskset -c 0-3 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 6 2>/dev/null | p "4 cores, chunks $c:"; done
