#!/bin/bash
# r02z: nfft 4096 STFT: which parts outside the FFT engine go on packed fp32x2 (KPK variants), and the complex-multiply operand order (lib_alt = -DNXS_CMUL_F1)
OUT=gpurun_out/r02z; mkdir -p $OUT
run_all() {
  for v in 9 10 0 11 12 13; do echo "NXS_STFT_VARIANT=$v (9 scalar, 10 engine only, 0 default = +split pass, 11 +window, 12 +window +conj stores, 13 engine + window)"; NXS_STFT_VARIANT=$v timeout 120 python tools/run_stft.py 128 60 4096 1024 10; done
  timeout 120 python tools/run_istft.py 32 60 1024 256 10
  timeout 200 python tools/run_fir.py 64 600 2049 3; timeout 200 python tools/run_fir.py 64 600 255 3
}
echo "== lib (swizzled operand first)" > $OUT/timings.txt; run_all >> $OUT/timings.txt 2>&1
cp nx_signal_b200/lib/libnxsignal_b200.so /tmp/main.so; cp nx_signal_b200/lib_alt/libnxsignal_b200.so nx_signal_b200/lib/libnxsignal_b200.so
echo "== lib_alt (NXS_CMUL_F1: broadcast first, MOV + FADD per product)" >> $OUT/timings.txt; run_all >> $OUT/timings.txt 2>&1
cp /tmp/main.so nx_signal_b200/lib/libnxsignal_b200.so
cat $OUT/timings.txt
