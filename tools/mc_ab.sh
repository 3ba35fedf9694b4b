#!/bin/bash
# A/B of marching-cubes scheduling variants at BASELINE config 4: kernel time from tools/mesh_bench.py for every library given.
for lib in "$@"; do
  echo "=== $lib"
  DIF_LIB_PATH=$lib timeout 300 python tools/mesh_bench.py 2>&1 | grep -E "^marching_cubes_kernel"
done
