mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_tail.log
rm -f gpurun_out/tune_pipe_c2.jsonl
for b in 16 1024 4096 8192 16384 32768; do timeout 300 python tools/tune_baked.py c2 $b 2>/dev/null | tee -a gpurun_out/tune_pipe_c2.jsonl; done
timeout 300 python tools/tune_baked.py c4 2>/dev/null | tee -a gpurun_out/tune_pipe_c2.jsonl
timeout 300 python tools/tune_baked.py c5 2>/dev/null | tee -a gpurun_out/tune_pipe_c2.jsonl
