#!/bin/bash
# usage: scratch/prof.sh <tag>   (runs under gpurun; outputs in gpurun_out/)
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in raster_bwd_kernel raster_fwd_kernel isect_emit_balanced_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K} -s 3 -c 1 -f -o gpurun_out/prof_${K}_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ref-cuda > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
