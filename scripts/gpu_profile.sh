#!/bin/bash
# ncu captures for profiles/ (run under gpurun, one GPU; gpurun_out/ must stay below 64 MiB).
# usage: scripts/gpu_profile.sh [config] [tag]
#   1. launch list of a short bench run (gpu__time_duration per launch)
#   2. DRAM bytes + duration of every BP sweep launch (roofline.traffic of a whole sweep)
#   3. --set full of the dominant kernels: bp4_kernel<4, 0> (class 4, non-first sweep), simscore3, planemap3
mkdir -p gpurun_out
CFG=${1:-c3}
TAG=${2:-r02}
B="python bench.py --config $CFG --steps 1 --warmup 1 --no-cpu --no-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_${CFG}_${TAG}.csv \
    $B > gpurun_out/bench_under_ncu_${CFG}_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:bp4_kernel -c 60 \
    --csv --log-file gpurun_out/bp_sweep_dram_${CFG}_${TAG}.csv $B > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:bp4_kernel<\(int\)4, \(bool\)0>' -s 1 -c 1 -f \
    -o gpurun_out/prof_bp4_${CFG}_${TAG} $B > gpurun_out/prof_bp4_${CFG}_${TAG}.log 2>&1
for K in ${KERNELS:-simscore3_kernel bp4_first_mapped_kernel depth3_kernel dda_codes_kernel}; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -f \
      -o gpurun_out/prof_${K}_${CFG}_${TAG} $B > gpurun_out/prof_${K}_${CFG}_${TAG}.log 2>&1
done
ls -la gpurun_out
