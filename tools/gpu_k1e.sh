#!/bin/bash
# K1 for one-tensor (elastic) tangents: the Gauss-sum-first kernel against the general one, timed at config B,
# register caps (TFEM_K1E_MINB resident CTAs), parity over the kernel/model suites, one full ncu capture.
mkdir -p gpurun_out
cd torch-fem_b200/csrc
for mb in 3 4; do
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -I../../include --expt-relaxed-constexpr -DTFEM_K1E_MINB=$mb -c integrate.cu -o integrate.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libtfem_b200.so error.o pattern.o integrate.o assemble.o krylov.o dcg.o residual.o amg.o
  echo "MINB=$mb"; (cd ../..; python tools/time_k1.py 150)
done
cd ../..
echo "TFEM_K1_ELASTIC=0"; TFEM_K1_ELASTIC=0 python tools/time_k1.py 150
python -m pytest tests/test_gpu_kernels.py tests/test_gpu_assembly.py tests/test_gpu_models.py -q -m gpu -x 2>&1 | tail -5
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_integrate_elastic -c 1 -o gpurun_out/k1e4 python tools/time_k1.py 150 > gpurun_out/k1e_ncu.log 2>&1
