#!/bin/bash
# ncu evidence for the tile kernel at HEAD (wind-aligned 240x60 tiles): launch list of one headline call + one full capture
cd "$(dirname "$0")/.."
TAG=r2t
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --launch-skip 900 -c 330 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-cascade --skip-layered --skip-cfg4 > $O/ncu_bench_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:tile_fly_kernel --launch-skip 40 -c 1 -o $O/tile_fly_$TAG -f \
    python tools/profile_target.py fly4096tile1 > $O/ncu_tile_fly_$TAG.log 2>&1
ls -la $O/*_$TAG*
