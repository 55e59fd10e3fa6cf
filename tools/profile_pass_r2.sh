#!/bin/bash
# Round 2: one GPU call that refreshes the evidence under gpurun_out/ (tools/make_profiles.py r2 turns it into
# profiles/). Numbers printed by runs under ncu / compute-sanitizer are never bench values.
set -x
TAG=${1:-r2}
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-cascade --skip-layered --skip-cfg4 > $O/ncu_bench_$TAG.log 2>&1
FULL="$NCU --set full --import-source on"
$FULL -k regex:colour_persistent_kernel --launch-skip 1 -c 1 -o $O/colour_persistent_$TAG -f python tools/profile_target.py fly4096 > $O/ncu_fly_$TAG.log 2>&1
$FULL -k regex:cascade_tile_kernel --launch-skip 2 -c 1 -o $O/cascade_sweep8192_$TAG -f python tools/profile_target.py cascade8192 > $O/ncu_sweep_$TAG.log 2>&1
$FULL -k regex:cascade_tile_kernel --launch-skip 2 -c 1 -o $O/cascade_sweep_at_rest_$TAG -f python tools/profile_target.py cascade8192relaxed >> $O/ncu_sweep_$TAG.log 2>&1
$FULL -k regex:cascade_tile_kernel --launch-skip 160 -c 1 -o $O/cascade_sweep_converged_$TAG -f python tools/profile_target.py cascade4096conv >> $O/ncu_sweep_$TAG.log 2>&1
# compute-sanitizer on HEAD: memcheck (global / shared addressing) and racecheck (shared-memory hazards) over the
# generate, sweep (dense-direct, dense-scan, pair-bitmap paths), colour-queue erode (persistent and launch-per-micro-round),
# lock-step erode and logical-stripe paths
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py all > $O/sanitizer_${tool}_$TAG.log 2>&1
  tail -4 $O/sanitizer_${tool}_$TAG.log
done
ls -la $O/*_$TAG.ncu-rep
