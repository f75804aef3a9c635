#!/bin/bash
# One GPU call: every variants/lib_<tag>.so x residency settings, headline and stage-2 raster shapes.
cd "$(dirname "$0")/.."
out=gpurun_out/variant_sweep.jsonl; : > $out
for lib in variants/lib_*.so; do
    tag=$(basename $lib .so); tag=${tag#lib_}
    R3DG_LIB_PATH=$PWD/$lib python tools/residency_sweep.py $tag 1000000 800 800 5 "${HEAD_SETTINGS:-0,0 6,4 5,4 4,3 3,3 4,2 3,2}" >> $out 2>> gpurun_out/variant_sweep.err
    R3DG_LIB_PATH=$PWD/$lib python tools/residency_sweep.py $tag 1500000 1600 1200 16 "${S2_SETTINGS:-0,0 4,3 3,3 3,2 2,2}" >> $out 2>> gpurun_out/variant_sweep.err
done
cat $out
