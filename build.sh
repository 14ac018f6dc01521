#!/bin/bash
# Build libnpdr_b200.so (sm_100a only) in-tree and the CPU oracle.  Usage: ./build.sh [-v]
set -e
cd "$(dirname "$0")"
EXTRA=""
[ "$1" = "-v" ] && EXTRA="-Xptxas -v"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $EXTRA -shared \
     -o npdr_b200/libnpdr_b200.so npdr_b200/csrc/api.cu npdr_b200/csrc/distance.cu npdr_b200/csrc/neighbors.cu \
     npdr_b200/csrc/regress.cu npdr_b200/csrc/multi.cu
make -s -C oracle
echo "built npdr_b200/libnpdr_b200.so and oracle/libnpdr_oracle.so"
