#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02k}
D2P_PARITY_LOG=gpurun_out timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -x -s -k "tensor_core_conv or c4 or vizdoom" 2>&1 | tail -25 | cut -c1-300
