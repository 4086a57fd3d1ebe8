#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r02f}
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 tools/dp_check.py > gpurun_out/${TAG}_dp_check_n2.json 2> gpurun_out/${TAG}_dp_check_n2.err
echo "dp_check rc=$?"; tail -3 gpurun_out/${TAG}_dp_check_n2.json; grep "^rank" gpurun_out/${TAG}_dp_check_n2.err
for mode in 1 0; do
D2P_DP_OVERLAP=$mode timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2965$mode tools/timeline.py > gpurun_out/${TAG}_timeline_n2_overlap$mode.txt 2>&1
echo "timeline overlap=$mode rc=$?"; grep " us " gpurun_out/${TAG}_timeline_n2_overlap$mode.txt
done
