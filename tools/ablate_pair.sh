#!/bin/bash
# GPU tool (development): the pair kernel under timing ablations (results are WRONG when CMHAR_ABLATE is set)
# bits: 1 no weight copies, 16 no attention MMAs, 32 no dense MMAs, 64 no epilogue work, 256 epilogue waits spin without back-off,
#       512 producer polls without back-off
for a in ${@:-0 768 113 881}; do
  echo "CMHAR_ABLATE=$a"; CMHAR_ABLATE=$a timeout 120 python tools/enc_pair_bench.py 65536 2>&1 | tail -1
done
