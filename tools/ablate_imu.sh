#!/bin/bash
# Development aid: in-situ timing ablations of the bf16 encoder kernel (results are wrong when a bit is set).
# bits: 1 no weight TMA after the first tile, 2 no epilogue smem stores, 4 no softmax math, 8 no LayerNorm,
#       16 no attention MMAs, 32 no dense MMAs
B=${1:-65536}
python tools/profile_imu.py $B bf16 4
python tools/profile_imu.py $B bf16 4 nohead
for m in 1 2 4 8 16 32 48 14 63; do CMHAR_ABLATE=$m python tools/profile_imu.py $B bf16 4 nohead; done
