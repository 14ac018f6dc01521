#!/bin/bash
# Build libnpdr_b200.so (sm_100a only) in-tree and the CPU oracle.  Usage: ./build.sh [-v]
# The translation units are compiled in parallel (objects under build/, git-ignored) and linked into one shared library.
set -e
cd "$(dirname "$0")"
EXTRA=""
[ "$1" = "-v" ] && EXTRA="-Xptxas -v"
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $EXTRA"
mkdir -p build
pids=()
for f in api distance neighbors kgrid regress regress_wide multi; do
    nvcc $FLAGS -c -o build/$f.o npdr_b200/csrc/$f.cu &
    pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o npdr_b200/libnpdr_b200.so build/api.o build/distance.o build/neighbors.o build/kgrid.o \
     build/regress.o build/regress_wide.o build/multi.o
make -s -C oracle
echo "built npdr_b200/libnpdr_b200.so and oracle/libnpdr_oracle.so"
