#!/bin/bash
# Development helper: builds libpbte_cuda.so with extra nvcc flags into tools/_variants/<name>/ (git-ignored, shipped to the GPU box).
#   tools/build_variant.sh <name> [-DPBTE_KC=3 ...]
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
PKG="$ROOT/dg-solver-for-pbte-with-mfem_b200"
name="$1"; shift
mkdir -p "$ROOT/tools/_variants/$name"
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" -I"$PKG/csrc" "$@" \
     -shared "$PKG/csrc/pbte_cuda.cu" -o "$ROOT/tools/_variants/$name/libpbte_cuda.so" -lnccl
cp "$PKG/lib/libpbte_host.so" "$ROOT/tools/_variants/$name/"
