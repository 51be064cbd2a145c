# Quick GPU pass: full GPU suite, smoke, bench lines of every BASELINE config (summary on stdout).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_tail.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py --no-cpu 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c1 --no-cpu 2>gpurun_out/bench_c1.err | tail -1 > gpurun_out/bench_c1.json
for wl in c3 c4 c5; do timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_$wl.err | tail -1 > gpurun_out/bench_$wl.json; done
python - <<'PY'
import json
for f in ("c2","c1","c3","c4","c5"):
    try:
        d=json.load(open(f"gpurun_out/bench_{f}.json")); r=d.get("roofline") or {}
        pb=(d.get("parity") or {}).get("forward")
        print(f, "ms/step",d["ms_per_step"],"value %.4g"%d["value"],"e2e %.4g"%d["e2e"]["value"],"fwd",r.get("kernel"),r.get("kernel_ms"),"fwd parity",json.dumps(pb)[:300])
    except Exception as e: print(f, "ERR", e)
PY
