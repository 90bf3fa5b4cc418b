set -x
for pf in 0 1; do echo "== L2_PREFETCH=$pf"; CMHAR_L2_PREFETCH=$pf timeout 300 python tools/bench_hbm_kernels.py 2>&1 | grep -E "maha_score|maha_accumulate|video_pool"; done > gpurun_out/pf_ab.txt 2>&1
for ln in 10 16 20 32; do echo "== lanes $ln"; timeout 300 python bench.py --steps 20 --warmup 5 --lanes $ln --no-cpu-baseline --no-sweep --workloads none 2>gpurun_out/lanes_$ln.err | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['timed_regions'], d['config'].get('step_latency_ms'))"; done > gpurun_out/lanes_ab.txt 2>&1
cat gpurun_out/pf_ab.txt gpurun_out/lanes_ab.txt
