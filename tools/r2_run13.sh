timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_sensor_gpu.py tests/test_trainer_gpu.py -x -q -k "per_env or environment or full_size or c3 or broad_phase or baked" 2>&1 | tail -5
timeout 600 python tools/tune_baked.py c3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['B'], {k:v for k,v in d.items() if isinstance(v,list)})"
timeout 600 python tools/tune_baked.py c3 8192 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['B'], {k:v for k,v in d.items() if isinstance(v,list)})"
timeout 600 python bench.py --workload c3 --steps 10 --warmup 3 --no-cpu 2>gpurun_out/b3.err | tail -1 > gpurun_out/bench_c3b.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c3b.json')); print('c3', d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['kernel_ms'], d['parity']['forward'], d['parity']['backward']['rel_err_vs_ball_gradient_norm'])"
tail -3 gpurun_out/b3.err
