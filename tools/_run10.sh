python tools/bench_fit.py 2>&1 | grep PF= | tee gpurun_out/fit_nsw16.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "maha" 2>&1 | tail -2
