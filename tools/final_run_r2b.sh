#!/bin/bash
# Round-2 (late) command list behind profiles/*_r2c*: GPU tests, the driver-style bench line, the reference arm, smoke.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu_r2c.log
cat gpurun_out/pytest_gpu_r2c.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -8 > gpurun_out/smoke_r2c.txt
cat gpurun_out/smoke_r2c.txt
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu_r2c_steps20.json 2> gpurun_out/bench_r2c.err
tail -2 gpurun_out/bench_r2c.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference_arm_r2c.json 2>> gpurun_out/bench_r2c.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_1gpu_r2c_steps20.json').read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"], "frames", (d["e2e"].get("from_frames") or {}).get("value"))
for k, v in d["workloads"].items():
    print(k, v.get("value"), v.get("ms_per_step"), v.get("breakdown_ms"))
print({k: round(v["frac"], 3) for k, v in d["roofline_scoring"].items()})
r = json.loads(open('gpurun_out/bench_reference_arm_r2c.json').read().strip().splitlines()[-1])
print("reference", r["value"], r["e2e"])
PY
