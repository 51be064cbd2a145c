mkdir -p gpurun_out
python tools/tune_baked.py c2 > gpurun_out/tune_baked_c2.json 2>gpurun_out/tune_baked_c2.err
python tools/tune_baked.py c1 > gpurun_out/tune_baked_c1.json 2>/dev/null
for wl in c3 c4 c5; do timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu 2>gpurun_out/bench_$wl.err | tail -1 > gpurun_out/bench_$wl.json; done
cat gpurun_out/tune_baked_c2.json gpurun_out/tune_baked_c1.json
