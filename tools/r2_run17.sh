timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "default_backward or sequential_backward or sliding or fused" 2>&1 | tail -2
for wl in c5 c2 c3; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --no-parity 2>/dev/null | tail -1 > gpurun_out/t.json; python -c "
import json; d=json.load(open('gpurun_out/t.json')); print('$wl', d['ms_per_step'], d['roofline']['kernel_ms'], round(d['ms_per_step']-d['roofline']['kernel_ms'],4))"; done
