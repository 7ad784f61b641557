#!/bin/bash
# register / spill report of one CUDA source: tools/ptxas_v.sh attn_sm100.cu [kernel-substring]
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I../../include -Xptxas -v -c $1 -o /tmp/$(basename $1 .cu).o 2>&1 | grep -A3 "${2:-Compiling}\|error" | grep -v "^--"
