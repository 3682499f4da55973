#!/bin/bash
# Build the library of an earlier commit into build/ab/libpsvae_old.so (for tools/ab_bench.sh).  usage: tools/build_old.sh <commit>
set -e
C=${1:-HEAD}
S=build/ab/src_old
rm -rf "$S" && mkdir -p "$S" build/ab
git archive "$C" pseudo_speaker_vae_b200/csrc include | tar -x -C "$S"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o build/ab/libpsvae_old.so "$S"/pseudo_speaker_vae_b200/csrc/psvae_b200.cu
echo "built build/ab/libpsvae_old.so from $C"
