ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c4.csv python tools/ncu_c4c5.py c4 > gpurun_out/ncu_c4.log 2>&1; tail -1 gpurun_out/ncu_c4.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_c5.csv python tools/ncu_c4c5.py c5 > gpurun_out/ncu_c5.log 2>&1; tail -1 gpurun_out/ncu_c5.log
