#!/bin/bash
# builds a TRACE copy of the library into scratch/ (run with GSAGE_LIB=scratch/libgsage_trace.so) (diagnostics only)
set -e
cd "$(dirname "$0")/.."
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -shared -Xptxas -O3 -DGS_TOP_TRACE -o scratch/libgsage_trace.so graphsage-pytorch_b200/csrc/*.cu
