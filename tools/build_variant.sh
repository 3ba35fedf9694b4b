#!/bin/bash
# Build an A/B variant of the library with extra nvcc defines:  tools/build_variant.sh <name> -DDIF_WAIT_MODE=2 ...
# -> tools/_build/libdifusion_b200_<name>.so ; run with DIF_LIB_PATH=tools/_build/libdifusion_b200_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/_build/obj_$name
for f in difusion_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-O3 "$@" -c $f -o tools/_build/obj_$name/$(basename $f .cu).o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/_build/libdifusion_b200_$name.so tools/_build/obj_$name/*.o
rm -rf tools/_build/obj_$name
echo tools/_build/libdifusion_b200_$name.so
