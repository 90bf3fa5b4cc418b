set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 2000 --warmup 10 > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r1z.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1z.json 2>> gpurun_out/bench_r1z.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep --lanes 1 > gpurun_out/plain_r1z.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1z.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep --lanes 1 > gpurun_out/ncu_r1z.log 2>&1; echo "launch list rc=$?"
python tools/bench_hbm_kernels.py > gpurun_out/hbm_r1z.txt 2>&1; cat gpurun_out/hbm_r1z.txt
