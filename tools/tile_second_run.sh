#!/bin/bash
# tile scheduler: sanitizers on small cases, the tile test file, timings
cd "$(dirname "$0")/.."
O=gpurun_out/tile_r2b
mkdir -p $O
for shape in 0 1; do
  for tool in memcheck racecheck; do
    TILE_SHAPE=$shape NO_ORACLE=1 timeout 400 compute-sanitizer --tool $tool --print-limit 5 python tools/colour_probe.py 250 0.25 -1 2 > $O/${tool}_shape$shape.log 2>&1
    echo "$tool shape $shape: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${tool}_shape$shape.log)"
  done
done
NO_ORACLE=1 timeout 300 python tools/colour_probe.py 4096 0.25 -1 2 noise 3 > $O/time4096.log 2>&1
NO_ORACLE=1 timeout 120 python tools/colour_probe.py 512 0.015625 -1 2 noise 3 > $O/time512_frame.log 2>&1
NO_ORACLE=1 timeout 300 python tools/colour_probe.py 1024 1.0 -1 2 flat 2 > $O/time1024_cfg2.log 2>&1
NO_ORACLE=1 TILE_SHAPE=1 timeout 300 python tools/colour_probe.py 1024 1.0 -1 2 flat 2 > $O/time1024_cfg2_shape1.log 2>&1
for f in time4096 time512_frame time1024_cfg2 time1024_cfg2_shape1; do echo "$f: $(grep -o 'erode [0-9.]* ms' $O/$f.log | tr '\n' ' ')"; done
timeout 1500 python -m pytest tests/test_gpu_tiles.py -x -q > $O/pytest_tiles.log 2>&1; tail -n 15 $O/pytest_tiles.log
