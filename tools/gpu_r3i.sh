#!/bin/bash
# r03i: does the achieved bandwidth depend on the footprint of one call?  nfft 4096 at 128 .. 1024 channels x 60 s (13 .. 106 GB), nfft 1024 at 8 .. 64 channels x 600 s (8 .. 66 GB)
OUT=gpurun_out/r03i; mkdir -p $OUT
{ for c in 128 256 512 768 1024; do timeout 300 python tools/run_stft.py $c 60 4096 1024 5; done
  for c in 8 16 32 64; do timeout 300 python tools/run_stft.py $c 600 1024 256 5; done
  nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
