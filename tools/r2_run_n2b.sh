mkdir -p gpurun_out
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_c2_n2.json
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2_n2.json"))
print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["plan"]["exchange"], d.get("exchange_self_check"))
for k,v in (d.get("sharded_workloads") or {}).items(): print(k, v["ms_per_step"], v["value"], v["scaling"])
print(d["parity"]["forward"])
PY
tail -n 3 gpurun_out/bench_n2.err
