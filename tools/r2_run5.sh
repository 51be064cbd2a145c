mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_tail.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c1 2>gpurun_out/bench_c1.err | tail -1 > gpurun_out/bench_c1.json
timeout 600 python bench.py --reference-regime --no-cpu 2>gpurun_out/bench_c2rr.err | tail -1 > gpurun_out/bench_c2_refregime.json
timeout 900 python bench.py --sharded --no-cpu --steps 20 2>gpurun_out/bench_sh.err | tail -1 > gpurun_out/bench_c2_sharded.json
python - <<'PY'
import json
for f in ("c2","c1","c2_refregime","c2_sharded"):
    try:
        d=json.load(open(f"gpurun_out/bench_{f}.json")); r=d["roofline"]
        print(f, "ms/step",d["ms_per_step"],"value",d["value"],"e2e",d["e2e"]["value"],"fwd",r["kernel"],r["kernel_ms"], d.get("sharded_workloads"))
    except Exception as e: print(f, "ERR", e)
PY
