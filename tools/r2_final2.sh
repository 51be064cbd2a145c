# Round-2 closing measurement pass on one GPU (everything lands in gpurun_out/, summaries are copied to profiles/ by hand).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_tail.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c1 2>gpurun_out/bench_c1.err | tail -1 > gpurun_out/bench_c1.json
timeout 600 python bench.py --reference-regime --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_c2_refregime.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_ref.json
for wl in c3 c4 c5; do timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_$wl.err | tail -1 > gpurun_out/bench_$wl.json; done
timeout 900 python bench.py --sharded --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_c2_sharded_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --no-parity > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv python bench.py --workload c5 --steps 2 --warmup 3 --no-cpu --no-graph --no-parity > gpurun_out/ncu_launch_c5.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ -c 6 -o gpurun_out/ncu_c2_final -f python tools/prof_fwd.py c2 default 0 2 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ --launch-skip 4 -c 4 -o gpurun_out/ncu_c5_final -f python tools/prof_fwd.py c5 default 131072 2 > gpurun_out/ncu_full_c5.log 2>&1
python - <<'PY'
import json
for f in ("c2","c1","c2_refregime","c3","c4","c5","ref","c2_sharded_n1"):
    try:
        d=json.load(open(f"gpurun_out/bench_{f}.json")); r=d.get("roofline") or {}
        print(f, "ms/step",d["ms_per_step"],"value %.4g"%d["value"],"e2e %.4g"%d["e2e"]["value"], d["e2e"].get("ms_per_step"),"fwd",r.get("kernel"),r.get("kernel_ms"),"frac",r.get("frac"), "launches", d.get("gpu_launches"))
    except Exception as e: print(f, "ERR", e)
PY
ls -la gpurun_out/*.ncu-rep
