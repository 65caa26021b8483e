#!/bin/bash
mkdir -p gpurun_out
for dbg in 8 9 10 11 15 27; do
RN_BP_DEBUG=$dbg timeout 300 python bench.py --no-cpu --no-e2e --steps 3 > gpurun_out/bench_dbg$dbg.log 2>&1
done
python - <<'PY'
import json
for i in (8,9,10,11,15,27):
    try:
        l=[x for x in open('gpurun_out/bench_dbg%d.log'%i) if x.startswith('{')][-1]; d=json.loads(l)
        print(i, d['ms_per_step'], d['stages_ms'], d['roofline']['launch_ms'])
    except Exception as e: print(i, 'fail', e, open('gpurun_out/bench_dbg%d.log'%i).read()[-2000:])
PY
