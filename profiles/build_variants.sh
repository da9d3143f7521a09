#!/bin/bash
# Builds A/B variants of librt795.so that differ only in -D flags of core.cu:  profiles/build_variants.sh name "-DX=1 -DY=2" ...
# Output: raytracer-795_b200/lib_ab/<name>/librt795.so (git-ignored); select with RT795_CORE_SO.
set -e
cd "$(dirname "$0")/../raytracer-795_b200"
make -s
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  mkdir -p lib_ab/$name
  ( nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr $flags -c csrc/core.cu -o lib_ab/$name/core.o 2> lib_ab/$name/build.log \
    && nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib_ab/$name/librt795.so lib_ab/$name/core.o build/bvh_device.o build/bvh_build.o build/api_error.o build/loader.o build/imageio.o -lz \
    && rm lib_ab/$name/core.o && echo "built $name" ) &
done
wait
