#!/bin/bash
# short device-resident bench summary (dev loop helper)
python bench.py --steps ${1:-10} --warmup 3 --no-cpu-baseline ${@:2} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'rank',round(d['ranking']['value']),'rank_ms',round(d['ranking']['ms_per_step'],3),'e2e',d['e2e'] and round(d['e2e']['value']))
print('stages',{k:round(v,3) for k,v in d['stage_ms_per_step'].items()})
print('gemm TF',round(d['roofline']['achieved'],1),'frontend GB/s',round(d['roofline_hbm']['achieved']), d['clocks'], 'launches', d['gpu_launches'])
"
