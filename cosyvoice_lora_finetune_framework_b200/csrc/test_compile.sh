#!/bin/bash
# quick syntax/ptxas check of one translation unit
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -c "$1" -o /tmp/$(basename "$1").o 2>&1 | grep -v "^$" | tail -${2:-40}
