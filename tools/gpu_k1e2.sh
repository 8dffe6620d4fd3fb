#!/bin/bash
# K1e: 8-byte against 16-byte broadcast loads of the tangent; full ncu capture of the solver-order assembly
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_assemble_solver' -c 1 \
   -o gpurun_out/k2s -f python tools/time_k1.py 150 > gpurun_out/k2s_ncu.log 2>&1
tail -2 gpurun_out/k2s_ncu.log
cd torch-fem_b200/csrc
for v in "-DTFEM_K1E_SCALAR_C=1" "-DTFEM_K1E_SCALAR_C=0"; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr $v -c integrate.cu -o integrate.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtfem_b200.so error.o pattern.o integrate.o assemble.o krylov.o dcg.o residual.o amg.o
  echo "== $v"; (cd ../..; python tools/time_k1.py 150 | cut -c1-40)
done
