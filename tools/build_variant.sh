#!/bin/bash
# Build a variant of libmdm_b200.so with extra nvcc defines, for A/B timing with tools/op_bench.py:
#   tools/build_variant.sh NAME [-DFOO=1 ...]   ->  variants/libmdm_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p variants/obj_$name
for f in motiondiffusion_moe_b200/csrc/*.cu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    -I include "$@" -c $f -o variants/obj_$name/$(basename $f .cu).o &
done
wait
nvcc -shared -o variants/libmdm_$name.so variants/obj_$name/*.o --cudart static
rm -rf variants/obj_$name
echo variants/libmdm_$name.so
