#!/bin/bash
# end-of-round profile set for C2: ncu launch list of two eager train steps, --set full of the fused Karel conv
# kernels, launch list of smoke()
mkdir -p gpurun_out
TAG=${TAG:-r03q}
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_2steps_c2.csv python tools/ncu_step.py 2 > gpurun_out/${TAG}_ncu_step.log 2>&1
echo "launch list rc=$?"; python tools/summarize_launches.py gpurun_out/${TAG}_launches_2steps_c2.csv half > gpurun_out/${TAG}_launches_2steps_c2.txt; head -14 gpurun_out/${TAG}_launches_2steps_c2.txt
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"karel_conv_(fwd|bwd)_fused" -s 2 -c 2 -o gpurun_out/${TAG}_karel_conv python tools/ncu_step.py 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/${TAG}_ncu_full.log
D2P_UNDER_NCU=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_smoke_launches.csv python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke_ncu.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke_ncu.log; python tools/summarize_launches.py gpurun_out/${TAG}_smoke_launches.csv 0 | head -5
ls -la gpurun_out | grep ${TAG}
