#!/bin/bash
# Timing experiments: build profiles/_exp/libvsl_<name>.so with extra nvcc flags for the two fused-loss translation
# units (the other objects come from the regular in-tree build).  usage: build_variant.sh <name> [nvcc flags...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p profiles/_exp
B=tf_depth_estimation_b200/build
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DVSL_DEV_V2_ONLY"
nvcc $F -c "$@" -o profiles/_exp/pair_$name.o tf_depth_estimation_b200/csrc/vsl_loss_pair.cu &
nvcc $F -c "$@" -Xptxas -v -o profiles/_exp/loss_$name.o tf_depth_estimation_b200/csrc/vsl_loss.cu 2>&1 | grep -A2 "loss_fused_kernelILi2ELb0ELb0" | grep -i "spill\|registers" &
wait
nvcc -shared -o profiles/_exp/libvsl_$name.so $B/vsl_ops.o profiles/_exp/loss_$name.o profiles/_exp/pair_$name.o $B/vsl_loss_ssim.o $B/vsl_flow.o $B/vsl_ext.o $B/vsl_optim.o
rm profiles/_exp/pair_$name.o profiles/_exp/loss_$name.o
echo built profiles/_exp/libvsl_$name.so
