mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "pipelined" 2>&1 | tail -5 | tee gpurun_out/pytest_pipe.log
rm -f gpurun_out/tune_pipe_c2.jsonl
for b in 16 1024 4096 8192 16384; do timeout 300 python tools/tune_baked.py c2 $b 2>/dev/null | tee -a gpurun_out/tune_pipe_c2.jsonl; done
