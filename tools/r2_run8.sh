mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_controller_gpu.py tests/test_trainer_gpu.py -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_ctl.log
