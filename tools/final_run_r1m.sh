set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 2000 --warmup 10 > gpurun_out/bench_r1m.json 2> gpurun_out/bench_r1m.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r1m.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r1m.json 2>> gpurun_out/bench_r1m.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep --lanes 1 > gpurun_out/plain_r1m.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1m.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-sweep --lanes 1 > gpurun_out/ncu_r1m.log 2>&1; echo "launch list rc=$?"
python tools/profile_imu.py 65536 bf16 3 nohead > gpurun_out/prof_plain_r1m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:imu_forward_bf16 -s 2 -c 1 -f -o gpurun_out/prof_imu_bf16_r1m python tools/profile_imu.py 65536 bf16 3 nohead > gpurun_out/ncu2_r1m.log 2>&1; echo "enc capture rc=$?"
python tools/profile_scoring.py > gpurun_out/prof_scoring_plain_r1m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:maha_score_tc -s 3 -c 1 -f -o gpurun_out/prof_maha_score_r1m python tools/profile_scoring.py > gpurun_out/ncu3_r1m.log 2>&1; echo "maha capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:logit_scores_ring -s 3 -c 1 -f -o gpurun_out/prof_logit_ring_r1m python tools/profile_scoring.py > gpurun_out/ncu4_r1m.log 2>&1; echo "logit capture rc=$?"
ls -la gpurun_out/*r1m*
