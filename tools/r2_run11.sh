timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_controller_gpu.py tests/test_trainer_gpu.py -x -q -k "fused or pipelined or trainer or default_backward or cached_host or device_inputs" 2>&1 | tail -8
timeout 300 python bench.py --no-cpu 2>gpurun_out/b.err | tail -1 > gpurun_out/bench_c2f.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c2f.json')); print('fused', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['plan']['backward_kernel']); print(d['parity']['backward']['rel_err_vs_ball_gradient_norm'], d['parity']['forward'])"
timeout 300 python bench.py --no-cpu --separate-launches 2>/dev/null | tail -1 > gpurun_out/bench_c2s.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c2s.json')); print('separate', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['gpu_launches'])"
timeout 300 python bench.py --no-cpu --workload c1 2>/dev/null | tail -1 > gpurun_out/bench_c1f.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c1f.json')); print('c1 fused', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'])"
tail -3 gpurun_out/b.err
