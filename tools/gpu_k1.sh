#!/bin/bash
cd torch-fem_b200/csrc
for v in 1 3 4 5 6; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr -DTFEM_K1_MINB=$v -Xptxas -v -c integrate.cu -o integrate.o 2> /tmp/ptxas_$v.log
  grep -A1 "k_integrateILi0ELi3ELi8ELi8ELi5ELb0" /tmp/ptxas_$v.log | grep -E "registers|spill" | head -2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtfem_b200.so error.o pattern.o integrate.o assemble.o krylov.o dcg.o residual.o amg.o
  echo "MINB=$v"; (cd ../..; python tools/time_k1.py 150)
done
