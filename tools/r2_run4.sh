mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_fwd -c 3 -o gpurun_out/ncu_c2_pipe8 -f python tools/prof_fwd.py c2 default 0 3 > gpurun_out/ncu_pipe.log 2>&1
tail -3 gpurun_out/ncu_pipe.log
