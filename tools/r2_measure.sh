# Round-2 measurement pass on one GPU: GPU tests, smoke, benches, ncu launch list + full captures.
mkdir -p gpurun_out
python -c "import jax" > gpurun_out/jax_probe.log 2>&1; echo "import jax exit $?" >> gpurun_out/jax_probe.log; ls baseline/_ref >> gpurun_out/jax_probe.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_tail.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 600 python bench.py 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/bench_c2.json
timeout 600 python bench.py --workload c1 2>gpurun_out/bench_c1.err | tail -1 > gpurun_out/bench_c1.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 2>gpurun_out/bench_ref.err | tail -1 > gpurun_out/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-graph --no-parity > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rollout_ -c 12 -o gpurun_out/ncu_c2_default -f python tools/prof_fwd.py c2 default 0 3 > gpurun_out/ncu_full.log 2>&1
for f in c2 c1 ref; do tail -c 3000 gpurun_out/bench_$f.json; echo; done
