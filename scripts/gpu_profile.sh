#!/bin/bash
# ncu captures for profiles/: launch list of one bench step + full-set capture of each hot kernel.
# usage: [KERNELS="bp4_kernel ..."] [SKIP=n] scripts/gpu_profile.sh [config] [tag]   (run under gpurun, one GPU;
# gpurun_out/ must stay below 64 MiB: two full-set reports per call)
mkdir -p gpurun_out
CFG=${1:-c3}
TAG=${2:-r01}
B="python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu --no-e2e"
if [ -z "$NOLIST" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${CFG}_${TAG}.csv \
    $B > gpurun_out/bench_under_ncu_${CFG}_${TAG}.log 2>&1
fi
for K in ${KERNELS:-bp4_kernel simmap3_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s ${SKIP:-2} -c 1 -f \
      -o gpurun_out/prof_${K}_${CFG}_${TAG} $B > gpurun_out/prof_${K}_${CFG}_${TAG}.log 2>&1
done
ls -la gpurun_out
