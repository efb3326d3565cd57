#!/bin/bash
# r02u: one-warp-per-frame ISTFT (radix 32 x 32, one exchange): parity + timings vs the default (T = 64, 16 x 8 x 8, XD)
OUT=gpurun_out/r02u; mkdir -p $OUT
for v in 0 2 3 4; do NXS_ISTFT_VARIANT=$v timeout 600 python -m pytest tests/test_istft_gpu.py -m gpu -q > $OUT/pytest_v$v.log 2>&1; echo "variant $v: $(tail -1 $OUT/pytest_v$v.log)"; done
{ for v in 0 2 3 4; do echo "NXS_ISTFT_VARIANT=$v (0 = default T=64 XD, 2 = warp/frame 384 thr, 3 = 256 thr, 4 = 320 thr)"; NXS_ISTFT_VARIANT=$v timeout 120 python tools/run_istft.py 32 60 1024 256 10; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
NXS_ISTFT_VARIANT=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:istft_rola -s 2 -c 1 -o $OUT/istft_w32_full -f python tools/run_istft.py 32 60 1024 256 2 > $OUT/ncu.log 2>&1; tail -1 $OUT/ncu.log
