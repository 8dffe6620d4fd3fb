RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 200 $RUN --master-port 29511 tools/multi_gpu_check.py --edge 24 2>&1 | grep -E "^nccl|^fused|^\{|Error" | head -20
echo ---- no interior
TFEM_DCG_NO_INTERIOR=1 timeout 200 $RUN --master-port 29512 tools/multi_gpu_check.py --edge 24 2>&1 | grep -E "^nccl|^fused|^\{|Error" | head -20
