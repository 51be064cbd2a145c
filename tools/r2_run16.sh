timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_bwd -c 2 -o gpurun_out/ncu_c5_bwd -f python tools/prof_fwd.py c5 default 131072 1 > gpurun_out/ncu_c5b.log 2>&1
tail -2 gpurun_out/ncu_c5b.log
