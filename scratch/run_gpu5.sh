#!/bin/bash
mkdir -p gpurun_out
for v in 4 8 16 32 64; do
RN_BP4_RPW=$v timeout 300 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/bench_rpw$v.log 2>&1
done
python - <<'PY'
import json
for i in (4,8,16,32,64):
    try:
        l=[x for x in open('gpurun_out/bench_rpw%d.log'%i) if x.startswith('{')][-1]; d=json.loads(l)
        print(i, d['ms_per_step'], d['stages_ms'], d['roofline']['launch_ms'])
    except Exception as e: print(i, 'fail', e)
PY
