#!/bin/bash
# dev-only: rebuild libmrt.so with a few compile-time variants on the GPU box and time cfg2
set -e
for v in "-DMRT_FWD_TPB=1" "-DMRT_FWD_TPB=2" "-DMRT_FWD_TPB=4"; do
  echo "=== $v"
  MRT_NVCC_EXTRA="$v" python -m mri_raytracer_b200.build --force > /dev/null
  MRT_NVCC_EXTRA="$v" python tools/quick_bench.py 4 1024 fold 2>&1 | grep -v "^$" | cut -c1-120
done
python -m mri_raytracer_b200.build --force > /dev/null
