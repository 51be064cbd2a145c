timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_tail.log
for b in 16 1024 4096; do timeout 120 python tools/tune_baked.py c2 $b 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['B'], {k:v[0] for k,v in d.items() if isinstance(v,list)})"; done
timeout 300 python bench.py --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_c2b.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c2b.json')); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['e2e']['value'])"
