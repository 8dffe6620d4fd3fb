#!/bin/bash
# K1 variants: RPT (row nodes per thread) x MINB (register cap via minimum resident CTAs), timed at config B.
cd torch-fem_b200/csrc
for cfg in "1 1" "1 4" "1 5" "2 1"; do
  set -- $cfg
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr -DTFEM_K1_RPT=$1 -DTFEM_K1_MINB=$2 -Xptxas -v -c integrate.cu -o integrate.o 2> /tmp/ptxas.log
  grep -A1 "k_integrateILi0ELi3ELi8ELi8ELi[0-9]*ELb0" /tmp/ptxas.log | grep -E "registers" | head -1
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtfem_b200.so error.o pattern.o integrate.o assemble.o krylov.o dcg.o residual.o amg.o
  echo "RPT=$1 MINB=$2"; (cd ../..; python tools/time_k1.py 150; python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "integrate or element_matrices or k0" 2>&1 | tail -1)
done
