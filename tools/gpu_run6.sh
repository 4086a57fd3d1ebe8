#!/bin/bash
# tensor-core convolution: A/B against the CUDA-core kernels, timing at C4, conv tests, launch list (tight time-outs)
mkdir -p gpurun_out
TAG=${TAG:-r02h}
timeout 150 python tools/conv_tc_check.py --time > gpurun_out/${TAG}_conv_tc_check.txt 2>&1
echo "rc=$?"; grep -n "^mode dx\|^==\|^C4" gpurun_out/${TAG}_conv_tc_check.txt | cut -c1-200
D2P_PARITY_LOG=gpurun_out timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -x -s -k "tensor_core_conv or c4 or vizdoom" 2>&1 | tail -8 | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_c4.csv python tools/ncu_c4c5.py c4 > gpurun_out/${TAG}_ncu_c4.log 2>&1
python tools/summarize_launches.py gpurun_out/${TAG}_launches_c4.csv half > gpurun_out/${TAG}_launches_c4_vizdoom.txt; head -16 gpurun_out/${TAG}_launches_c4_vizdoom.txt
