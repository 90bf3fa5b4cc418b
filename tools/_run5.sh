timeout 900 python -m pytest tests/test_gpu_maha_finalize.py tests/test_gpu_parity.py tests/test_gpu_pipeline.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/fin_tests.txt
cat gpurun_out/fin_tests.txt
python tools/bench_maha_finalize.py 2>&1 | grep -v Warn | tee gpurun_out/fin_bench.txt
