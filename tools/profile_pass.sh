#!/bin/bash
# One GPU call that refreshes the evidence under gpurun_out/ (tools/make_profiles.py turns it into profiles/).
# Numbers printed by runs under ncu are never bench values.
set -x
TAG=${1:-r1}
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 2200 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-cascade --skip-layered > $O/ncu_bench.log 2>&1
FULL="$NCU --set full --import-source on"
$FULL -k regex:par_round_kernel --launch-skip 660 -c 1 -o $O/par_round_$TAG -f python tools/profile_target.py fly4096 > $O/ncu_par.log 2>&1
$FULL -k regex:cascade_tile_stamped_kernel --launch-skip 660 -c 1 -o $O/cascade_stamped_$TAG -f python tools/profile_target.py fly4096 >> $O/ncu_par.log 2>&1
$FULL -k regex:cascade_tile_kernel --launch-skip 2 -c 1 -o $O/cascade_sweep8192_$TAG -f python tools/profile_target.py cascade8192 > $O/ncu_sweep.log 2>&1
$FULL -k regex:cascade_tile_kernel --launch-skip 2 -c 1 -o $O/cascade_sweep_at_rest_$TAG -f python tools/profile_target.py cascade8192relaxed >> $O/ncu_sweep.log 2>&1
$FULL -k regex:cascade_tile_kernel --launch-skip 160 -c 1 -o $O/cascade_sweep_converged_$TAG -f python tools/profile_target.py cascade4096conv >> $O/ncu_sweep.log 2>&1
$FULL -k regex:noise_stack_kernel -c 1 -o $O/noise_stack_$TAG -f python tools/profile_target.py fly4096 >> $O/ncu_par.log 2>&1
ls -la $O/*_$TAG.ncu-rep
