#!/usr/bin/env bash
# TEST INFRASTRUCTURE: builds the CPU oracle (plain C + OpenMP).  -ffp-contract=off is required:
# the FMA association is pinned explicitly with fmaf() (see oracle_raster.c header).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
gcc -O2 -fopenmp -ffp-contract=off -fno-fast-math -fPIC -shared -Wall -Wno-unknown-pragmas \
    "$HERE/oracle_raster.c" "$HERE/oracle_bvh.c" -o "$HERE/liboracle_raster.so" -lm
echo "[oracle] built $HERE/liboracle_raster.so"
