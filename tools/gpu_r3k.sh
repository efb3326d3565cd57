#!/bin/bash
# r03k: burst against sustained (power-capped) kernel times, packed fp32x2 plans against the scalar ones
OUT=gpurun_out/r03k; mkdir -p $OUT
{
echo "== STFT cfg2 (scalar default)"; timeout 100 python tools/run_sustained.py stft 8 600 1024 256 3
echo "== STFT nfft 4096, 128 ch: packed (default), scalar (variant 9), engine-only packed (variant 10)"
timeout 100 python tools/run_sustained.py stft 128 60 4096 1024 3; NXS_STFT_VARIANT=9 timeout 100 python tools/run_sustained.py stft 128 60 4096 1024 3; NXS_STFT_VARIANT=10 timeout 100 python tools/run_sustained.py stft 128 60 4096 1024 3
echo "== ISTFT cfg5: packed warp-per-frame (default), scalar warp-per-frame (NXS_ISTFT_SCALAR), scalar T=64 (NXS_ISTFT_SCALAR + variant 2)"
timeout 100 python tools/run_sustained.py istft 32 60 1024 256 3; NXS_ISTFT_SCALAR=1 timeout 100 python tools/run_sustained.py istft 32 60 1024 256 3; NXS_ISTFT_SCALAR=1 NXS_ISTFT_VARIANT=2 timeout 100 python tools/run_sustained.py istft 32 60 1024 256 3
echo "== FIR cfg4: packed (default), scalar (variant 8)"
timeout 100 python tools/run_sustained.py fir 64 600 2049 0 3; NXS_FIR_VARIANT=8 timeout 100 python tools/run_sustained.py fir 64 600 2049 0 3
echo "== STFT nfft 2048, 64 ch: packed engine (default), scalar (variant 9)"
timeout 100 python tools/run_sustained.py stft 64 60 2048 512 3; NXS_STFT_VARIANT=9 timeout 100 python tools/run_sustained.py stft 64 60 2048 512 3
} > $OUT/sustained.txt 2>&1; cat $OUT/sustained.txt
