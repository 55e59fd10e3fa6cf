#!/bin/bash
# first GPU runs of the tile-resident scheduler: memcheck on a tiny case, parity at 512^2, timing at the headline size
cd "$(dirname "$0")/.."
O=gpurun_out/tile_r2a
mkdir -p $O
NO_ORACLE=1 timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/colour_probe.py 192 0.25 -1 2 > $O/memcheck.log 2>&1; echo "memcheck rc $?" >> $O/memcheck.log
timeout 300 python tools/colour_probe.py 512 0.25 -1 2 > $O/parity512.log 2>&1
NO_ORACLE=1 timeout 300 python tools/colour_probe.py 4096 0.25 -1 2 noise 3 > $O/time4096.log 2>&1
NO_ORACLE=1 TILE_ROUNDS=8 timeout 300 python tools/colour_probe.py 4096 0.25 -1 2 noise 2 > $O/time4096_k8.log 2>&1
NO_ORACLE=1 timeout 120 python tools/colour_probe.py 512 0.015625 -1 2 noise 3 > $O/time512_frame.log 2>&1
NO_ORACLE=1 timeout 300 python tools/colour_probe.py 1024 1.0 -1 2 flat 2 > $O/time1024_cfg2.log 2>&1
tail -n 4 $O/*.log
