#!/bin/bash
# dev-only: rebuild libmrt.so with a few compile-time variants on the GPU box and time cfg2
set -e
for v in "$@"; do
  echo "=== $v"
  MRT_NVCC_EXTRA="$v" python -m mri_raytracer_b200.build --force > /dev/null
  MRT_NVCC_EXTRA="$v" python bench.py --no-cpu --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.1f G/s  ms/step %.3f  march launch %.3f ms'%(d['value']/1e9,d['ms_per_step'],d['roofline']['avg_launch_ms']))"
  MRT_NVCC_EXTRA="$v" python tools/quick_bench.py 4 1024 fold 2>&1 | grep '"skip": 1' | cut -c1-120
done
python -m mri_raytracer_b200.build --force > /dev/null
