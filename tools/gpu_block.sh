#!/bin/bash
python -m pytest tests/test_gpu_amg.py -q -m gpu -x -k "block_vcycle" 2>&1 | tail -3
python -m pytest tests/test_gpu_models.py -q -m gpu -x -k "Modal or modal" 2>&1 | tail -3
python tools/time_vcycle_block.py 150 2>&1 | tail -4
