mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_peer_gpu.py -q 2>&1 | tail -3 | tee gpurun_out/pytest_peer_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 2>gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_c2_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/trainer_multi_gpu_check.py 2>gpurun_out/trainer_n2.err | tail -1
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c2_n2.json"))
print("N=2 ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d["plan"]["exchange"], d.get("exchange_self_check"))
for k,v in (d.get("sharded_workloads") or {}).items(): print(k, v["ms_per_step"], v["value"], v["scaling"])
PY
tail -3 gpurun_out/bench_n2.err gpurun_out/trainer_n2.err
