#!/bin/bash
mkdir -p gpurun_out
./scratch/mb_smem > gpurun_out/mb_smem.log 2>&1
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
RN_BP_IMPL=3 timeout 300 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/bench_impl3.log 2>&1
cat gpurun_out/mb_smem.log
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for i in (3,):
    try:
        l=[x for x in open('gpurun_out/bench_impl%d.log'%i) if x.startswith('{')][-1]; d=json.loads(l)
        print(i, d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['launch_ms'], d['roofline']['frac'])
    except Exception as e: print(i, 'fail', e, open('gpurun_out/bench_impl%d.log'%i).read()[-2000:])
PY
