#!/bin/bash
# ncu evidence for the tile-resident scheduler: launch list of one headline call + one full capture of a mid-call fly kernel
# usage: profile_tile.sh TAG [LIB] [K]
TAG=${1:-r2t}
export SWE_B200_LIB=${2:-simplewinderosion_b200/libswe_b200.so}
export TILE_ROUNDS=${3:-16}
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --launch-skip 20 -c 520 --csv --log-file $O/launches_tile_$TAG.csv \
    python tools/profile_target.py fly4096tile1 > $O/ncu_tile_list_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:tile_fly_kernel --launch-skip 40 -c 1 -o $O/tile_fly_$TAG -f \
    python tools/profile_target.py fly4096tile1 > $O/ncu_tile_fly_$TAG.log 2>&1
ls -la $O/*_$TAG*
