#!/bin/bash
# Round-2 GPU check (run under gpurun, one GPU): GPU test suite, smoke, a short bench line.
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40) > gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
(timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3) > gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench_c3.log 2>&1
tail -c 6000 gpurun_out/bench_c3.log
