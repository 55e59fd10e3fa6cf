#!/bin/bash
# final evidence of the round for the tile scheduler at HEAD: sanitizers, bench line (1 GPU)
cd "$(dirname "$0")/.."
O=gpurun_out/final_r2t
mkdir -p $O
for shape in 0 1 2; do
  for tool in memcheck racecheck; do
    TILE_SHAPE=$shape NO_ORACLE=1 timeout 400 compute-sanitizer --tool $tool --print-limit 5 python tools/colour_probe.py 250 0.25 -1 2 > $O/${tool}_shape$shape.log 2>&1
    echo "$tool shape $shape: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $O/${tool}_shape$shape.log)"
  done
done
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_1gpu.json 2> $O/bench_1gpu.err; tail -c 300 $O/bench_1gpu.err
python -c "
import json; d=json.load(open('$O/bench_1gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['layered_extension']['ms_per_step'], d['dune_formation_cfg2']['ms_per_step'], d['cfg4_strong']['ms_per_step'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2>> $O/bench_1gpu.err; head -c 400 $O/bench_reference.json
