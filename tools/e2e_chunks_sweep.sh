#!/bin/bash
# e2e of the default workload for several pipeline shapes: chunk weights of compress_batch / decompress_batch
p() { python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value'],1), 'MPix/s device', round(d['e2e']['value'],1), 'e2e', round(d['e2e']['s_per_step']*1e3,1), 'ms')"; }
run() { CRDR_PIPELINE_WEIGHTS_COMPRESS=$1 CRDR_PIPELINE_WEIGHTS=$2 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 8 2>/dev/null | p "compress weights [$1], decompress weights [$2]:"; }
run "" ""
run "1,2" "2,1"
run "1,1" "2,1"
run "1,2" "1,1"
run "2,3" "3,2"
run "1,2" "3,1"
