#!/bin/bash
# r03u: ISTFT nfft 128 on the register-overlap-add kernel with half-warp groups against the gather kernel: parity, racecheck, timings
OUT=gpurun_out/r03u; mkdir -p $OUT
timeout 900 python -m pytest tests/test_istft_gpu.py tests/test_istft_c2r_gpu.py -m gpu -q > $OUT/pytest.log 2>&1; echo "istft: $(tail -1 $OUT/pytest.log)"
cat > /tmp/rc128.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
rng = np.random.default_rng(0)
for hop in (64, 32, 16):
    z = (rng.standard_normal((3, 700, 128)) + 1j * rng.standard_normal((3, 700, 128))).astype(np.complex64); w = o.hann(128)
    y = nx.istft(torch.from_numpy(z).cuda(), torch.from_numpy(w).cuda(), overlap_length=128 - hop, fft_length=128)
    yo = o.istft_fast(z, w, overlap_length=128 - hop, fft_length=128)
    e = np.abs(y.cpu().numpy() - yo).max() / np.abs(yo).max(); print("istft 128 /", hop, "rel err %.2e" % e); assert e < 1e-4
PY
timeout 600 compute-sanitizer --tool racecheck python /tmp/rc128.py > $OUT/racecheck.log 2>&1; tail -3 $OUT/racecheck.log
timeout 600 compute-sanitizer --tool memcheck python /tmp/rc128.py > $OUT/memcheck.log 2>&1; tail -2 $OUT/memcheck.log
{ for shape in "64 60 128 32" "64 60 128 64" "64 60 128 16"; do echo "ISTFT $shape: register overlap-add (default), gather kernel (NXS_ISTFT_NO_ROLA128)"; timeout 120 python tools/run_istft.py $shape 10; NXS_ISTFT_NO_ROLA128=1 timeout 120 python tools/run_istft.py $shape 10; done; } > $OUT/timings.txt 2>&1; cat $OUT/timings.txt
