#!/bin/bash
# rebuild with a few load-batching settings of the fold kernels and time them (run on the GPU box)
for cfg in "3 5" "1 3" "2 3" "1 2" "2 9"; do
  set -- $cfg
  MRT_NVCC_EXTRA="-DMRT_FOLD_SL1=$1 -DMRT_FOLD_G4=$2" python -c "from mri_raytracer_b200 import build; build.build()" || exit 1
  echo "SL1=$1 G4=$2 $(MRT_NVCC_EXTRA="-DMRT_FOLD_SL1=$1 -DMRT_FOLD_G4=$2" python tools/time_fold.py 2>&1 | tail -1)"
done
