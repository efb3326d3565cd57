#!/bin/bash
# r03h: ISTFT cfg5: warp-per-frame with the carry in shared memory (12 / 10 warps per SM) on packed fp32x2 against the register-carry default (8 warps)
OUT=gpurun_out/r03h; mkdir -p $OUT
timeout 900 python -m pytest tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
{ for v in 0 5 6 4 0 5; do echo "NXS_ISTFT_VARIANT=$v (0 default: registers, 256 thr; 5: smem carry 384 thr; 6: smem carry 320 thr; 4: registers 320 thr)"; NXS_ISTFT_VARIANT=$v timeout 120 python tools/run_istft.py 32 60 1024 256 10; done
  timeout 120 python tools/run_istft_c2r.py 32 60 1024 256 10; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
