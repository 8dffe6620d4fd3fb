N=${1:-2}
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 200 $RUN --master-port 29511 tools/multi_gpu_check.py --edge 6 --hexa2 2>&1 | grep -E "^\{|Error" | head
timeout 600 $RUN --master-port 29513 bench.py --gpus $N --config C --edge 64 --steps 2 --warmup 2 2>gpurun_out/r1h_bench_c_n$N.err > gpurun_out/r1h_bench_c_n$N.json
cat gpurun_out/r1h_bench_c_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], d['config']['cg_iterations'], d['config']['per_iteration_ms'], d['e2e']['value'], d['roofline'])"
tail -3 gpurun_out/r1h_bench_c_n$N.err
timeout 600 python bench.py --config C --edge 48 --steps 2 --warmup 2 2>&1 | tail -2
