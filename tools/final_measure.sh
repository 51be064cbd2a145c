# End-of-round check on one GPU: full GPU test suite, smoke, the tpc-kernel benches (C2, C2 reference regime, C1).
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_tail.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py 2>gpurun_out/f_bench_c2.err | tail -1 > gpurun_out/f_bench_c2.json
python bench.py --reference-regime 2>/dev/null | tail -1 > gpurun_out/f_bench_c2_refregime.json
python bench.py --workload c1 2>/dev/null | tail -1 > gpurun_out/f_bench_c1.json
for f in c2 c2_refregime c1; do python -c "
import json
d=json.load(open('gpurun_out/f_bench_$f.json')); r=d['roofline']; print('$f', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], r['kernel'], r['kernel_ms'], 'frac', r['frac'], 'cpu', (d.get('cpu_baseline') or {}).get('value'), d['clocks'])"; done
