#!/bin/bash
# Last call of the round on 2 GPUs: what the driver runs (tests, smoke, both arms at N = 1) and the N = 2 launch of both arms.
TAG=${1:-r2val2}
bash tools/gpu_r2_validate.sh $TAG
bash tools/gpu_r2_driver_like.sh ${TAG}_n2 2
