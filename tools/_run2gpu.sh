timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2> gpurun_out/multi_r2c.log | tail -3
grep "multi rank" gpurun_out/multi_r2c.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_r2c_steps20.json 2> gpurun_out/bench_2gpu_r2c.err
tail -2 gpurun_out/bench_2gpu_r2c.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_2gpu_r2c_steps20.json').read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "frames", (d["e2e"].get("from_frames") or {}).get("value"))
for k, v in d["workloads"].items():
    print(k, v.get("value"), v.get("ms_per_step"), v.get("breakdown_ms"), v.get("collective"))
PY
