#!/bin/bash
# Round-end measurement set (one gpurun call): GPU tests, smoke, every bench workload, the reference arm, launch list.
set -x
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > $O/final_gputest.log 2>&1; tail -3 $O/final_gputest.log
python __graft_entry__.py --smoke > $O/final_smoke.log 2>&1; tail -2 $O/final_smoke.log
timeout 600 python bench.py > $O/final_bench.json 2> $O/final_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/final_bench_reference.json 2>> $O/final_bench.err
for w in kodak1 clic uhd train train_gan; do timeout 600 python bench.py --workload $w --steps 10 --warmup 5 --no-cpu-baseline > $O/final_bench_$w.json 2>> $O/final_bench.err; done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1600 --csv --log-file $O/final_launches.csv python tools/one_step.py 24 > $O/final_ncu.log 2>&1
for f in $O/final_bench*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1])
print('$f', round(d['value'],2), d['unit'], round(d['ms_per_step'],2), 'ms', 'e2e', round(d['e2e']['value'],2), 'frac', round(d.get('roofline',{}).get('frac',0),4))"; done
