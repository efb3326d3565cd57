#!/bin/bash
# r03d: FIR real-packed kernel, store loop: predicated constant-offset stores (lib) against per-element branches (lib_alt = -DNXS_FIR_OLD_STORES)
OUT=gpurun_out/r03d; mkdir -p $OUT
timeout 900 python -m pytest tests/test_fir_conv_gpu.py tests/test_host_pipeline_gpu.py tests/test_full_size_gpu.py -m gpu -q -k "fir or conv or cfg4" > $OUT/pytest_fir.log 2>&1; tail -2 $OUT/pytest_fir.log
run_all() { timeout 200 python tools/run_fir.py 64 600 2049 3; timeout 200 python tools/run_fir.py 64 600 513 3; timeout 200 python tools/run_fir.py 64 600 2049 3; }
echo "== lib (predicated stores)" > $OUT/timings.txt; run_all >> $OUT/timings.txt 2>&1
cp nx_signal_b200/lib/libnxsignal_b200.so /tmp/main.so; cp nx_signal_b200/lib_alt/libnxsignal_b200.so nx_signal_b200/lib/libnxsignal_b200.so
echo "== lib_alt (per-element branches)" >> $OUT/timings.txt; run_all >> $OUT/timings.txt 2>&1
cp /tmp/main.so nx_signal_b200/lib/libnxsignal_b200.so
cat $OUT/timings.txt
