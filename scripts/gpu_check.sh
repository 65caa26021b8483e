#!/bin/bash
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/pytest_gpu.log
for impl in ${IMPLS:-4 3}; do
RN_BP_IMPL=$impl timeout 300 python bench.py --no-cpu --no-e2e --steps 5 > gpurun_out/bench_impl$impl.log 2>&1
done
tail -4 gpurun_out/pytest_gpu.log
python - <<'PY'
import json,os
for i in os.environ.get('IMPLS','4 3').split():
    try:
        l=[x for x in open('gpurun_out/bench_impl%s.log'%i) if x.startswith('{')][-1]; d=json.loads(l)
        print(i, d['value'], d['ms_per_step'], d['stages_ms'], d['roofline']['launch_ms'], d['roofline']['frac'])
    except Exception as e: print(i, 'fail', e, open('gpurun_out/bench_impl%s.log'%i).read()[-2000:])
PY
