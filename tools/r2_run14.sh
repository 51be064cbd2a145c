timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -c 1 -o gpurun_out/ncu_c5_fwd -f python tools/prof_fwd.py c5 default 131072 2 > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log
