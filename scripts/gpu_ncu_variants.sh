#!/bin/bash
# ncu --set full of one kernel (KERNEL regex, COUNT launches after SKIP) for each library variant in VARIANTS.
mkdir -p gpurun_out
cp raynet_b200/libraynet_b200.so /tmp/lib_main.so
for v in ${VARIANTS}; do
  cp raynet_b200/variants/$v.so raynet_b200/libraynet_b200.so
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KERNEL}" -s ${SKIP:-2} -c ${COUNT:-1} -f -o gpurun_out/ncu_${v} \
      python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 ${BENCH_ARGS:-} > gpurun_out/ncu_${v}.log 2>&1
  tail -2 gpurun_out/ncu_${v}.log | cut -c1-200
done
cp /tmp/lib_main.so raynet_b200/libraynet_b200.so
