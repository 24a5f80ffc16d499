#!/bin/bash
# One GPU session (runs under gpurun): parity tests, bench lines, per-kernel times of a training iteration, ncu evidence.
# usage: profiles/tools/gpu_session.sh <tag> [full]      (full: also the complete bench line)
TAG=${1:-r2}; MODE=${2:-quick}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -rs -s > $O/${TAG}_gputest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_gputest.log
timeout 120 python bench.py --quick > $O/${TAG}_benchB.json 2> $O/${TAG}_benchB.err
timeout 120 python bench.py --quick --config D > $O/${TAG}_benchD.json 2> /dev/null
timeout 120 python profiles/tools/train_prof.py --config B > $O/${TAG}_trainB.json 2> $O/${TAG}_trainB.err
timeout 120 python profiles/tools/train_prof.py --config D > $O/${TAG}_trainD.json 2> /dev/null
if [ "$MODE" = full ]; then timeout 900 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; fi
# launch list of two steps (shares of the step; cold-cache, serialised)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --quick > $O/bench_under_ncu_${TAG}.log 2>&1
for K in raster_bwd_kernel raster_fwd_kernel tile_bin_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K} -s 4 -c 2 -f -o $O/prof_${K}_${TAG} \
    python bench.py --steps 1 --warmup 1 --quick > $O/ncu_${K}_${TAG}.log 2>&1
done
# fastgs path: front / back kernels, the EWA instantiations of the blend kernels, the filtered walks
timeout 120 python profiles/tools/fastgs_prof.py --config B > $O/${TAG}_fastgsB.json 2> $O/${TAG}_fastgsB.err
timeout 120 python profiles/tools/fastgs_prof.py --config D > $O/${TAG}_fastgsD.json 2> /dev/null
for K in fgs_front_kernel fgs_back_kernel raster_bwd_kernel raster_fwd_kernel tile_bin_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K} -s 2 -c 1 -f -o $O/prof_fastgs_${K}_${TAG} \
    python profiles/tools/fastgs_prof.py --iters 1 > $O/ncu_fastgs_${K}_${TAG}.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_fastgs_${TAG}.csv \
    python profiles/tools/fastgs_prof.py --iters 2 > $O/fastgs_under_ncu_${TAG}.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_fastgs_ref_${TAG}.csv \
    python profiles/tools/fastgs_prof.py --iters 2 --reference > $O/fastgs_ref_under_ncu_${TAG}.log 2>&1
for K in ssim_l1_kernel fused_front_kernel fused_back_kernel; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:${K} -s 2 -c 1 -f -o $O/prof_${K}_${TAG} \
    python profiles/tools/train_prof.py --iters 1 > $O/ncu_${K}_${TAG}.log 2>&1
done
tail -5 $O/${TAG}_gputest.log
