timeout 600 python -m pytest tests/test_parity_gpu.py tests/test_controller_gpu.py -x -q -k "pipelined or fused" 2>&1 | tail -3
timeout 600 python tools/bench_trainer.py 2>/dev/null | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fused.csv python tools/fused_iter.py 4096 2 > gpurun_out/ncu_fused.log 2>&1
tail -2 gpurun_out/ncu_fused.log
