#!/bin/bash
# r02n: FIR mid-size K on the real-packed N = 4096 kernel (dispatch fixed); parity
OUT=gpurun_out/r02n; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fir_conv_gpu.py tests/test_mel_gpu.py tests/test_golden_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
{ for k in 131 193 255 385 513; do for v in 3 0; do echo "K=$k variant $v (3 = pair kernel, 0 = default)"; NXS_FIR_VARIANT=$v timeout 200 python tools/run_fir.py 64 600 $k 3; done; done
timeout 200 python tools/run_mel.py 32 60 1024 256 128 0 48000; timeout 200 python tools/run_mel.py 8 600 1024 256 128 0 48000; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
