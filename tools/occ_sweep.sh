#!/usr/bin/env bash
# Occupancy sweep of the compositors on the GPU box: rebuilds the two compositor objects with different
# __launch_bounds__ minimum-CTA settings (R3DG_{FWD,BWD}_OCC{2,5}) and prints the stage times of the headline and
# stage-2 benches.  Usage (under gpurun): bash tools/occ_sweep.sh > gpurun_out/occ_sweep.txt
set -u
cd "$(dirname "$0")/.."
run() {
  export R3DG_NVCC_DEFS="$1"
  touch relightable3dgaussian_b200/csrc/composite.cu relightable3dgaussian_b200/csrc/composite_bwd.cu
  python -m relightable3dgaussian_b200.build >/dev/null 2>&1 || { echo "$1: build failed"; return; }
  python bench.py --steps 40 --warmup 10 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['stage_ms']; print('$1 | headline', round(d['value'],1), 'fwd', round(s['composite_fwd'],4), 'bwd', round(s['composite_bwd'],4))"
  python bench.py --workload stage2 --steps 10 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); s=d['raster_stage_ms']; print('$1 | stage2  ', round(d['ms_per_step'],3), 'fwd', round(s['composite_fwd'],4), 'bwd', round(s['composite_bwd'],4))"
}
run "-DR3DG_FWD_OCC2=8 -DR3DG_BWD_OCC2=6 -DR3DG_FWD_OCC5=5 -DR3DG_BWD_OCC5=4"
run "-DR3DG_FWD_OCC2=9 -DR3DG_BWD_OCC2=7 -DR3DG_FWD_OCC5=6 -DR3DG_BWD_OCC5=5"
run "-DR3DG_FWD_OCC2=7 -DR3DG_BWD_OCC2=5 -DR3DG_FWD_OCC5=4 -DR3DG_BWD_OCC5=3"
run "-DR3DG_FWD_OCC2=10 -DR3DG_BWD_OCC2=8 -DR3DG_FWD_OCC5=7 -DR3DG_BWD_OCC5=4"
run "-DR3DG_FWD_OCC2=6 -DR3DG_BWD_OCC2=4 -DR3DG_FWD_OCC5=3 -DR3DG_BWD_OCC5=2"
# restore the default build
unset R3DG_NVCC_DEFS
touch relightable3dgaussian_b200/csrc/composite.cu relightable3dgaussian_b200/csrc/composite_bwd.cu
python -m relightable3dgaussian_b200.build >/dev/null 2>&1
