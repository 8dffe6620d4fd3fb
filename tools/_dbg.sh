N=${1:-4}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 200 $RUN --master-port 29511 tools/multi_gpu_check.py --edge 6 --hexa2 2>&1 | grep -E "^\{|Error" | head
timeout 800 $RUN --master-port 29513 bench.py --gpus $N --config C --steps 2 --warmup 2 2>gpurun_out/r1k_bench_c_n$N.err > gpurun_out/r1k_bench_c_n$N.json
cat gpurun_out/r1k_bench_c_n$N.json
tail -3 gpurun_out/r1k_bench_c_n$N.err
