#!/bin/bash
# Builds an alternative libmod16_b200 with extra -D flags for kernel A/B experiments.
# usage: scripts/build_variant.sh <name> <-Dflags...>   -> mod16_b200/lib/variants/lib<name>.so
name=$1; shift
mkdir -p mod16_b200/lib/variants
cd mod16_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr \
  -Xcompiler -fPIC --fmad=true -shared "$@" -o ../lib/variants/lib$name.so \
  *.cu
