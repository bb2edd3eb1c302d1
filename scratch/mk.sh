#!/bin/bash
# usage: scratch/mk.sh <name> [-D...]   builds scratch/<name> from trainbench.cu + the product kernels
cd "$(dirname "$0")"; CS=../object-detection-collection-pytorch_b200/csrc; name=$1; shift
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -I ../include -I $CS "$@" -o $name trainbench.cu $CS/yh_api.cu $CS/yh_train.cu $CS/yh_nms.cu 2>&1 | grep -E "error|Error" ; echo built $name
