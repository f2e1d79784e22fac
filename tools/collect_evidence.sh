#!/bin/bash
# Round-2 (second session) evidence run on a B200 box:  gpurun --timeout 1500 -- 'bash tools/collect_evidence.sh [tag]'
# Writes ncu captures / launch list / sanitizer logs into gpurun_out/ (copied from there into profiles/ by hand after
# reading them with tools/ncu_summary.py).  Nothing printed by a run under ncu or the sanitizer is a bench value.
TAG=${1:-r02b}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
# 1. launch list of the bench step (share of every kernel in the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 > $OUT/${TAG}_launches_bench.log 2>&1
# 2. full captures of exactly the launches bench.py's roofline probes time
$NCU -k regex:nn_scan_f32_kernel -s 2 -c 1 -o $OUT/${TAG}_full_nn_scan python tools/prof_nn_bench.py > $OUT/${TAG}_full_nn_scan.log 2>&1
$NCU -k regex:gn_accum_staged -s 5 -c 1 -o $OUT/${TAG}_full_gn_accum python tools/sweep_gn.py --child 256 > $OUT/${TAG}_full_gn_accum.log 2>&1
$NCU -k regex:bwd_points_staged -s 20 -c 1 -o $OUT/${TAG}_full_bwd_112 python tools/prof_bwd.py > $OUT/${TAG}_full_bwd_112.log 2>&1
$NCU -k regex:bwd_points_staged -s 290 -c 1 -o $OUT/${TAG}_full_bwd_124 python tools/prof_bwd.py > $OUT/${TAG}_full_bwd_124.log 2>&1
$NCU -k regex:bwd_pair_kernel -s 20 -c 1 -o $OUT/${TAG}_full_bwd_pair python tools/prof_bwd.py > $OUT/${TAG}_full_bwd_pair.log 2>&1
NO_TORCH_PROFILER=1 $NCU -k regex:gn_update_kernel -s 12 -c 1 -o $OUT/${TAG}_full_gn_update python tools/step_profile.py 256 brute > $OUT/${TAG}_full_gn_update.log 2>&1
# gpurun brings back at most 64 MiB: summarise the captures here and keep only the two that the source view is read from
for r in nn_scan gn_accum bwd_112 bwd_124 bwd_pair gn_update; do
  python tools/ncu_summary.py raw $OUT/${TAG}_full_$r.ncu-rep > $OUT/${TAG}_full_$r.md 2>/dev/null
done
rm -f $OUT/${TAG}_full_bwd_112.ncu-rep $OUT/${TAG}_full_bwd_pair.ncu-rep $OUT/${TAG}_full_gn_update.ncu-rep $OUT/${TAG}_full_nn_scan.ncu-rep
# 3. sanitizers over the tests that drive the kernels changed in this session (staged accumulate / reverse kernels with and
#    without ragged batches, PDL-chained reverse loop, unrolled solve, sync-free forward, split batches)
SEL='fixture or dense or ragged or sync_free or more_pairs or n_valid or zero or weight'
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "$SEL" > $OUT/${TAG}_memcheck.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "tile_ring or planar or subnormal or every_weight or early_stop or far_outliers" >> $OUT/${TAG}_memcheck.log 2>&1
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "tile_ring or planar" > $OUT/${TAG}_racecheck.log 2>&1
compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "tile_ring and pt2pl" > $OUT/${TAG}_synccheck.log 2>&1
echo done
