mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -x -q -k "pipelined" 2>&1 | tail -5 | tee gpurun_out/pytest_pipe.log
rm -f gpurun_out/tune_pipe_c2.jsonl
for b in 16 1024 4096 8192; do timeout 120 python tools/tune_baked.py c2 $b 2>/dev/null | tee -a gpurun_out/tune_pipe_c2.jsonl; done
timeout 120 python tools/tune_baked.py c1 2>/dev/null | tee -a gpurun_out/tune_pipe_c2.jsonl
timeout 120 python tools/pipe_timers.py run c2 4096 16 | tee gpurun_out/pt16.json
timeout 120 python tools/pipe_timers.py run c2 4096 16 0 | tee gpurun_out/pt16rr.json
