#!/bin/bash
# Round 2, tile-resident scheduler: one GPU call that refreshes the evidence under gpurun_out/ (tools/make_profiles.py r2t
# turns it into profiles/). Numbers printed by runs under ncu / compute-sanitizer are never bench values.
cd "$(dirname "$0")/.."
TAG=r2t
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
# launch list of ONE timed headline call (three warm-up calls of ~300 launches each are skipped)
$NCU --metrics gpu__time_duration.sum --launch-skip 900 -c 330 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --skip-cpu --skip-cascade --skip-layered --skip-cfg4 > $O/ncu_bench_$TAG.log 2>&1
$NCU --set full --import-source on -k regex:tile_fly_kernel --launch-skip 40 -c 1 -o $O/tile_fly_$TAG -f \
    python tools/profile_target.py fly4096tile1 > $O/ncu_tile_fly_$TAG.log 2>&1
# statistical parity of the tile schedule against the sequential oracle (same cases as r2_stat_parity_density.txt)
P=$O/stat_parity_tiles_$TAG.txt
echo "# tools/colour_probe.py on B200 (round 2): parallel mode, TILE-RESIDENT scheduler (schedule 2, automatic stagger, automatic tile geometry unless TILE_SHAPE is given) vs the sequential oracle" > $P
echo "# columns: n density stagger schedule map calls" >> $P
probe() { echo "## $*" >> $P; timeout 600 python tools/colour_probe.py "$@" 2>&1 | grep "^call" >> $P; }
probe 1024 0.25 -1 2 noise 1
echo "## (TILE_SHAPE=2: 240x60 tiles, long side along the wind -- the geometry the 4096^2 headline picks)" >> $P
TILE_SHAPE=2 probe 1024 0.25 -1 2 noise 1
echo "## (TILE_SHAPE=1: 120x120 tiles)" >> $P
TILE_SHAPE=1 probe 1024 0.25 -1 2 noise 1
probe 512 0.25 -1 2 noise 2
probe 256 1.0 -1 2 flat 1
probe 256 1.0 -1 2 noise 2
probe 512 1.0 -1 2 noise 1
probe 1024 0.03 -1 2 noise 1
cat $P | cut -c1-400
ls -la $O/*_$TAG*
