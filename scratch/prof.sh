#!/bin/bash
# usage: scratch/prof.sh <tag>   (runs under gpurun; outputs in gpurun_out/)
TAG=${1:-r1}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in raster_bwd_kernel raster_fwd_kernel; do
ncu --set full --clock-control none --import-source on -k regex:${K} -s 3 -c 1 -f -o gpurun_out/prof_${K}_${TAG} \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
