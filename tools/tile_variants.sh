#!/bin/bash
# A/B of tile-scheduler build variants (tools/build_variant.py) on the headline call
cd "$(dirname "$0")/.."
O=gpurun_out/tile_variants
mkdir -p $O
run() {  # name lib
  SWE_B200_LIB=$2 NO_ORACLE=1 timeout 120 python tools/colour_probe.py 4096 0.25 -1 2 noise 3 > $O/$1.log 2>&1
  echo "$1: $(grep -o 'erode [0-9.]* ms' $O/$1.log | tr '\n' ' ') $(grep -o 'micro-rounds [0-9]*' $O/$1.log | tail -1)"
}
L=simplewinderosion_b200
for v in "$@"; do run $v $L/libswe_$v.so; done
