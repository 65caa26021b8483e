#!/bin/bash
# A/B runs of library variants (raynet_b200/variants/*.so, built by raynet_b200.build.build_variant):
# every variant is copied over libraynet_b200.so and benched with the same short command.
mkdir -p gpurun_out
cp raynet_b200/libraynet_b200.so /tmp/lib_main.so
for v in ${VARIANTS}; do
  cp raynet_b200/variants/$v.so raynet_b200/libraynet_b200.so
  timeout 300 python bench.py --no-cpu --no-e2e --steps ${STEPS:-8} ${BENCH_ARGS:-} > gpurun_out/variant_$v.log 2>&1
  if [ -n "${CHECK:-}" ]; then timeout 600 python -m pytest tests -x -q -m gpu -k "${CHECK}" > gpurun_out/variant_${v}_check.log 2>&1; echo "$v check: $(tail -1 gpurun_out/variant_${v}_check.log)"; fi
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    l = [x for x in open('gpurun_out/variant_%s.log' % v) if x.startswith('{')][-1]; d = json.loads(l)
    print(v, 'ms/step %.2f' % d['ms_per_step'], {k: (round(x, 3) if x is not None else x) for k, x in d['stages_ms'].items()}, 'frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'])
except Exception as e:
    print(v, 'fail', e, open('gpurun_out/variant_%s.log' % v).read()[-1500:])
PY
done
cp /tmp/lib_main.so raynet_b200/libraynet_b200.so
