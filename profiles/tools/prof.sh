#!/bin/bash
# usage: profiles/tools/prof.sh <tag> [kernel regexes...]   (runs under gpurun; outputs in gpurun_out/)
# 1. launch list of two bench steps (gpu__time_duration per launch, cold-cache, serialised: compare SHARES)
# 2. one `--set full` capture per kernel regex (default: the blend kernels and the intersect walk)
TAG=${1:-r2}; shift
KERNELS=${@:-"raster_bwd_kernel raster_fwd_kernel tile_bin_kernel"}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --quick > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in $KERNELS; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K} -s 4 -c 2 -f -o gpurun_out/prof_${K}_${TAG} \
    python bench.py --steps 1 --warmup 1 --quick > gpurun_out/ncu_${K}_${TAG}.log 2>&1
done
