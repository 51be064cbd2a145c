mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sensor_gpu.py tests/test_trainer_gpu.py tests/test_parity_gpu.py -x -q -k "sensor or environment or float64 or larger or trainer or closest or inner or per_env" 2>&1 | tail -15 | tee gpurun_out/pytest_sensor.log
