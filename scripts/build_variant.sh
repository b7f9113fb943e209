#!/bin/bash
# builds an experimental variant of the library next to the main one: scripts/build_variant.sh NAME -DFLAG=... ; use with
# KLAMPT_B200_LIB=$PWD/klampt_b200/_variants/libklampt_b200_NAME.so
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/var_$name klampt_b200/_variants
for f in kb_kernels kb_engine kb_lbvh kb_closest kb_raycast; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Wno-deprecated-gpu-targets "$@" -c klampt_b200/csrc/$f.cu -o build/var_$name/$f.o &
done
wait
nvcc -shared -Wno-deprecated-gpu-targets -o klampt_b200/_variants/libklampt_b200_$name.so build/var_$name/kb_kernels.o build/var_$name/kb_engine.o build/var_$name/kb_lbvh.o build/var_$name/kb_closest.o build/var_$name/kb_raycast.o
echo built klampt_b200/_variants/libklampt_b200_$name.so
