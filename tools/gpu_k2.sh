#!/bin/bash
# K2 variants: contributions in flight per lane, timed at config B.
cd torch-fem_b200/csrc
for v in 2 4 8; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr -DTFEM_K2_INFLIGHT=$v -Xptxas -v -c assemble.cu -o assemble.o 2> /tmp/ptxas.log
  grep -A1 "k_assembleILi3ELi8E" /tmp/ptxas.log | grep -E "registers" | head -1
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtfem_b200.so error.o pattern.o integrate.o assemble.o krylov.o dcg.o residual.o amg.o
  echo "INFLIGHT=$v"; (cd ../..; python tools/time_k1.py 150; python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "assemble" 2>&1 | tail -1)
done
