#!/bin/bash
# usage: build_variant.sh <name> [-DFLAG ...]  ->  tools/abtest/lib_<name>.so (A/B measurements; same flags as r2p2_b200.build())
set -e
cd "$(dirname "$0")/../.."
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared "$@" \
  -o tools/abtest/lib_$name.so r2p2_b200/csrc/r2p2_capi.cu 2>&1 | grep -v "warning #550\|r2p2_handle h = nullptr\|\^$\|^$\|Remark" || true
echo built $name
