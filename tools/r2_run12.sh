timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_controller_gpu.py -x -q -k "fused" 2>&1 | tail -3
for b in 16 256 1024; do for mode in "" "--separate-launches"; do timeout 300 python bench.py --no-cpu --no-parity --balls $b $mode 2>/dev/null | tail -1 > gpurun_out/t.json; python -c "
import json; d=json.load(open('gpurun_out/t.json')); print($b, '$mode', d['ms_per_step'], d['e2e']['ms_per_step'], d['gpu_launches'])"; done; done
timeout 300 python bench.py --no-cpu --workload c1 2>/dev/null | tail -1 > gpurun_out/bench_c1f.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c1f.json')); print('c1', d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'])"
