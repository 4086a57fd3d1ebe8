#!/bin/bash
# ncu --set full of the tensor-core convolution kernels in one C4 step (no source import: 12 launches)
mkdir -p gpurun_out
TAG=${TAG:-r02j}
timeout 900 ncu --set full --clock-control none -k regex:"conv_tc_(gather|dw)_kernel" -c 12 -o gpurun_out/${TAG}_conv_tc python tools/ncu_c4c5.py c4 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/${TAG}_ncu_full.log; ls -la gpurun_out/
