#!/bin/sh
# libiic_b200.so with csrc/finish.cu compiled -DIIC_FIN_TRACE (tools/fin_trace.py)
set -e
cd "$(dirname "$0")/.."
P=mi-based-regularized-semi-supervised-segmentation_b200
mkdir -p tools/_bin
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DIIC_FIN_TRACE -I include -c $P/csrc/finish.cu -o tools/_bin/finish_trace.o
OBJS=$(ls $P/csrc/_obj/*.o | grep -v "/finish.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o tools/_bin/libiic_fintrace.so $OBJS tools/_bin/finish_trace.o
