#!/bin/bash
# Multi-GPU check + exchange microbenchmark (run under gpurun --gpus N): N=${N:-2}
mkdir -p gpurun_out
N=${N:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ -z "${SKIP_CHECK:-}" ]; then
timeout 600 $T tests/multi_gpu_check.py > gpurun_out/multi_check_n$N.log 2>&1
grep -a "MULTI_GPU_CHECK" gpurun_out/multi_check_n$N.log | cut -c1-1200 || tail -30 gpurun_out/multi_check_n$N.log
grep -a -A25 "MULTI_GPU_CHECK failed" gpurun_out/multi_check_n$N.log | head -60
fi
if [ -z "${SKIP_EXCHANGE:-}" ]; then
timeout 300 $T scratch/exchange_bench.py > gpurun_out/exchange_n$N.log 2>&1
grep -a "^{" gpurun_out/exchange_n$N.log || tail -30 gpurun_out/exchange_n$N.log
fi
if [ -n "${BENCH:-}" ]; then
  for C in ${BENCH}; do
    timeout 600 $T bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --collective $C ${BENCH_ARGS:-} > gpurun_out/bench_n${N}_$C.log 2>&1
    grep -a "^{" gpurun_out/bench_n${N}_$C.log | tail -1 > gpurun_out/bench_n${N}_$C.json
    python - gpurun_out/bench_n${N}_$C.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g' % d['value'], 'ms %.3f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], {k: (round(x, 3) if x is not None else x) for k, x in d['stages_ms'].items()}, d['config'].get('collective', '')[:90])
except Exception as e:
    print('fail', e); print(open(sys.argv[1].replace('.json', '.log')).read()[-3000:])
PY
  done
fi
