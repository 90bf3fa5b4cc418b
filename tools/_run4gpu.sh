timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_4gpu_r2c_steps20.json 2> gpurun_out/bench_4gpu_r2c.err
tail -2 gpurun_out/bench_4gpu_r2c.err | cut -c1-300
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_4gpu_r2c_steps20.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "frames", (d["e2e"].get("from_frames") or {}))
for k, v in d["workloads"].items():
    print(k, v.get("value"), v.get("ms_per_step"), v.get("breakdown_ms"), v.get("exchange", {}).get("ms") if isinstance(v.get("exchange"), dict) else None)
PY
