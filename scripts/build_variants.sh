#!/bin/bash
# Builds of the same sources with other ptxas flags for k_fused_sm's file (the kernel sits at the 168-register cap, where the
# schedule ptxas finds moves the kernel time by several per cent): wumingpic2d_b200/variants/lib_<name>.so, selected with WM_LIB.
#   usage: bash scripts/build_variants.sh   (after python -m wumingpic2d_b200.build)
set -e
cd "$(dirname "$0")/.."
mkdir -p wumingpic2d_b200/variants /tmp/wmvar
OBJ=wumingpic2d_b200/_obj
build() {
  name=$1; shift
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" -c wumingpic2d_b200/csrc/fused5_kernel.cu -o /tmp/wmvar/f5_$name.o
  objs=$(ls $OBJ/*.o | grep -v fused5_kernel.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o wumingpic2d_b200/variants/lib_$name.so /tmp/wmvar/f5_$name.o $objs -lnccl -lcudart
}
build r0 -Xptxas --register-usage-level=0 &
build r7 -Xptxas --register-usage-level=7 &
build r10 -Xptxas --register-usage-level=10 &
wait
ls -la wumingpic2d_b200/variants
