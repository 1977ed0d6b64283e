#!/bin/bash
# Diagnostic: build a tuning variant of the library (extra -D macros) into align3r_b200/lib/var_<name>.so and print the
# instruction mix of the folded flow scan's chunk loop.   tools/variant.sh <name> [-DMACRO[=v] ...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=align3r_b200/lib/var_$name.so
src="align3r_b200/csrc/a3r_common.cu align3r_b200/csrc/ga_kernels.cu align3r_b200/csrc/ga_handle.cu align3r_b200/csrc/mst_kernels.cu align3r_b200/csrc/motion_kernels.cu"
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared \
    -I include -I align3r_b200/csrc "$@" -Xptxas -v -o $out $src 2> /tmp/variant_$name.ptxas
grep -A1 "k_accumulate_v4ILb1ELb0ELb0ELb1" /tmp/variant_$name.ptxas | grep -E "registers|spill" | head -3
A3R_LIB=$out python tools/sass_loop.py k_accumulate_v4ILb1ELb0ELb0ELb1
