#!/bin/bash
# build/libpsa_<name>.so: the product library with extra -D switches (kernel tuning experiments; PSA_LIB_PATH selects it)
set -e
cd "$(dirname "$0")/../rust-pseudoaligner_b200/csrc"
name=$1; shift
mkdir -p ../../build
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-deprecated-declarations \
  --expt-relaxed-constexpr -diag-suppress 20012 "$@" -shared -o ../../build/libpsa_$name.so psa_api.cu process_reads.cpp -ldl -lz
