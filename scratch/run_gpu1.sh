#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
for impl in 3 2; do
  RN_BP_IMPL=$impl timeout 300 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/bench_impl$impl.log 2>&1
done
(./scratch/mb_l1 148 1024 200000 | grep -E "RED |blocks"; ./scratch/mb_l1 37 1024 200000 | grep -E "RED |blocks") > gpurun_out/mb_sm.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for i in (3,2):
    try:
        l=[x for x in open('gpurun_out/bench_impl%d.log'%i) if x.startswith('{')][-1]; d=json.loads(l)
        print(i, d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['launch_ms'], d['roofline']['frac'])
    except Exception as e: print(i, 'fail', e, open('gpurun_out/bench_impl%d.log'%i).read()[-2000:])
PY
cat gpurun_out/mb_sm.log
