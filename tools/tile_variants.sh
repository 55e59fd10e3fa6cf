#!/bin/bash
# A/B of tile-scheduler build variants (tools/build_variant.py) on the headline call
cd "$(dirname "$0")/.."
O=gpurun_out/tile_variants
mkdir -p $O
run() {  # name lib K
  SWE_B200_LIB=$2 TILE_ROUNDS=$3 NO_ORACLE=1 timeout 120 python tools/colour_probe.py 4096 0.25 -1 2 noise 2 > $O/$1.log 2>&1
  echo "$1: $(grep -o 'erode [0-9.]* ms' $O/$1.log | tr '\n' ' ') $(grep -o 'micro-rounds [0-9]*' $O/$1.log | tail -1)"
}
L=simplewinderosion_b200
run v0_k16 $L/libswe_b200.so 16
run v1_k16 $L/libswe_v1.so 16
run v2_k16 $L/libswe_v2.so 16
run v3_k24 $L/libswe_v3.so 24
run v4_k32 $L/libswe_v4.so 32
run v5_k24 $L/libswe_v5.so 24
run v6_k32 $L/libswe_v6.so 32
run v7_k12 $L/libswe_v7.so 12
run v8_k24 $L/libswe_v8.so 24
run v3_k16 $L/libswe_v3.so 16
run v4_k24 $L/libswe_v4.so 24
