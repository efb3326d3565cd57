#!/bin/bash
# (historical: this run used a global NXS_F32X2 build switch; the choice is per plan now -- Plan::PK, NXS_STFT_VARIANT / NXS_ISTFT_SCALAR / NXS_FIR_VARIANT=8)
# r02x: packed fp32x2 complex arithmetic (FADD2/FMUL2/FFMA2) in the FFT engine against the scalar build: parity + timings
OUT=gpurun_out/r02x; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_stft_gpu.py tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py tests/test_fir_conv_gpu.py tests/test_mel_gpu.py tests/test_golden_gpu.py tests/test_host_pipeline_gpu.py -m gpu -q -x > $OUT/pytest.log 2>&1; tail -2 $OUT/pytest.log
run_all() {
  timeout 120 python tools/run_stft.py 8 600 1024 256 10; timeout 120 python tools/run_stft.py 128 60 4096 1024 10
  timeout 120 python tools/run_stft.py 64 60 2048 512 10; timeout 120 python tools/run_stft.py 32 60 8192 2048 10
  timeout 120 python tools/run_istft.py 32 60 1024 256 10; timeout 120 python tools/run_istft.py 32 60 2048 512 10
  timeout 120 python tools/run_istft_c2r.py 32 60 1024 256 10
  timeout 200 python tools/run_fir.py 64 600 2049 3; timeout 200 python tools/run_fir.py 64 600 255 3
  timeout 120 python tools/run_mel.py 2>/dev/null | tail -4
}
echo "== packed (NXS_F32X2=1)" > $OUT/timings.txt; run_all >> $OUT/timings.txt 2>&1
cp nx_signal_b200/lib/libnxsignal_b200.so /tmp/packed.so; cp nx_signal_b200/lib_scalar/libnxsignal_b200.so nx_signal_b200/lib/libnxsignal_b200.so
echo "== scalar (NXS_F32X2=0)" >> $OUT/timings.txt; run_all >> $OUT/timings.txt 2>&1
cp /tmp/packed.so nx_signal_b200/lib/libnxsignal_b200.so
echo "== packed again" >> $OUT/timings.txt; run_all >> $OUT/timings.txt 2>&1
cat $OUT/timings.txt
