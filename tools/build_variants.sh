#!/bin/bash
# Build kernel variants of the library side by side (variants/lib_<tag>.so, git-ignored) for one-call A/B runs on the
# GPU box: tools/build_variants.sh "tag1:-DR3DG_FWD_ILP=1" "tag2:-DR3DG_FWD_ILP=4 -DR3DG_FWD_OCC2=6" ...
# The in-tree library is rebuilt with the default flags at the end.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
for spec in "$@"; do
    tag="${spec%%:*}"; defs="${spec#*:}"
    R3DG_NVCC_DEFS="$defs" python -m relightable3dgaussian_b200.build --force > /dev/null
    cp relightable3dgaussian_b200/libr3dg_b200.so "variants/lib_${tag}.so"
    echo "built variants/lib_${tag}.so  ($defs)"
done
python -m relightable3dgaussian_b200.build --force > /dev/null
