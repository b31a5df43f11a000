#!/bin/bash
# Build an A/B variant of libgeobi.so with extra -D flags on feast_tcagg.cu:  profiles/build_variant.sh NAME -DTCAGG_AB=2 ...
# -> build_variants/libgeobi_NAME.so (git-ignored, travels to the GPU box; load it with GEOBI_LIB_PATH=...).
set -e
name=$1; shift
cd "$(dirname "$0")/../geobi_gnn_b200/csrc"
make -s all
mkdir -p build/var ../../build_variants
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
  -I../../include -I. "$@" -Xptxas -v -c feast_tcagg.cu -o build/var/feast_tcagg_$name.o 2> build/var/$name.ptxas.log
objs=$(ls build/*.o | grep -v feast_tcagg.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build_variants/libgeobi_$name.so $objs build/var/feast_tcagg_$name.o -cudart static
grep -A3 "feast_tcagg_64_32_kernel" build/var/$name.ptxas.log | grep -i "registers\|spill" | head -6
