mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 50 --warmup 5 2>gpurun_out/bench_n8.err | tail -1 > gpurun_out/bench_c2_n8.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2_n8.json"))
print("N=8 ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["plan"]["exchange"][:40], d.get("exchange_self_check"))
for k,v in (d.get("sharded_workloads") or {}).items(): print(k, v["ms_per_step"], v["value"], v["scaling"])
PY
tail -n 3 gpurun_out/bench_n8.err
