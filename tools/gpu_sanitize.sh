# compute-sanitizer over small cases of every hot kernel: memcheck (OOB / misaligned) and racecheck (shared-memory hazards)
OUT=gpurun_out/sanitize; mkdir -p $OUT
cat > /tmp/san_cases.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, ".")
import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
rng = np.random.default_rng(0)
def chk(name, got, want, tol=1e-5):
    got = np.asarray(got.cpu() if hasattr(got, "cpu") else got); e = np.abs(got - want).max() / max(np.abs(want).max(), 1e-30)
    print(name, "rel err %.2e" % e, "OK" if e <= tol else "FAIL"); assert e <= tol
for nfft, hop, pad in [(1024, 256, "valid"), (1024, 250, "reflect"), (1024, 441, "same"), (2048, 512, "valid"), (4096, 1024, "valid"), (512, 128, "valid"), (8192, 2048, "valid"), (256, 64, "valid")]:
    x = rng.standard_normal((2, 12 * nfft)).astype(np.float32); w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=48000, window_padding=pad)
    z, _, _ = nx.stft(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), **kw)
    zo, _, _ = o.stft_fast(x, w, **kw); chk(f"stft {nfft}/{hop}/{pad}", torch.view_as_real(z), np.stack([zo.real, zo.imag], -1))
    if nfft <= 4096:
        m = nx.stft_mel(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), mel_bins=40, **kw)
        mo = np.stack([o.stft_to_mel(zo[c], 48000, nfft, 40) for c in range(2)]); chk(f"stft_mel {nfft}", m, mo)
for nfft, hop in [(1024, 256), (1024, 512), (1024, 128), (1024, 250), (1024, 1024), (512, 128), (2048, 512), (4096, 1024), (256, 100), (2048, 700), (4096, 1000), (128, 32)]:
    z = (rng.standard_normal((2, 40, nfft)) + 1j * rng.standard_normal((2, 40, nfft))).astype(np.complex64); w = o.hann(nfft)
    y = nx.istft(torch.from_numpy(z).cuda(), torch.from_numpy(w).cuda(), overlap_length=nfft - hop, fft_length=nfft)
    yo = o.istft_fast(z, w, overlap_length=nfft - hop, fft_length=nfft); chk(f"istft {nfft}/{hop}", torch.view_as_real(y), np.stack([yo.real, yo.imag], -1))
for K, L in [(2049, 30000), (255, 20000), (4097, 30000), (65, 5000), (600, 20001)]:
    x = rng.standard_normal((2, L)).astype(np.float32); taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    y = nx.convolution.convolve(torch.from_numpy(x).cuda(), torch.from_numpy(taps).cuda()[None, :], mode="same", method="fft")
    from scipy.signal import oaconvolve
    full = oaconvolve(x.astype(np.float64), taps.astype(np.float64)[None, :], mode="full", axes=-1); s = (K - 1) // 2
    chk(f"fir K={K}", y, full[:, s:s + L].astype(np.float32))
z = (rng.standard_normal((2, 50, 1024)) + 1j * rng.standard_normal((2, 50, 1024))).astype(np.complex64)
m = nx.stft_to_mel(torch.from_numpy(z).cuda(), 48000, fft_length=1024, mel_bins=128)
chk("stft_to_mel", m, np.stack([o.stft_to_mel(z[c], 48000, 1024, 128) for c in range(2)]))
# fused log-mel at a sampling rate whose filters cover every bin (all threads of the bin-major epilogue active)
for nfft, hop, mels in [(1024, 256, 80), (512, 128, 128), (2048, 512, 64)]:
    x = rng.standard_normal((2, 12 * nfft)).astype(np.float32); w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, sampling_rate=16000)
    m = nx.stft_mel(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), mel_bins=mels, **kw)
    zo, _, _ = o.stft_fast(x, w, **kw)
    chk(f"stft_mel16k {nfft}/{mels}", m, np.stack([o.stft_to_mel(zo[c], 16000, nfft, mels) for c in range(2)]))
# c2r ISTFT: packed kernels (tight odd row length: alternating 8-byte row alignment), extension path
for nfft, hop in [(1024, 256), (1024, 512), (1024, 128), (512, 128), (2048, 512), (4096, 1024), (1024, 250), (256, 64)]:
    K = nfft // 2 + 1
    z1 = (rng.standard_normal((2, 40, K)) + 1j * rng.standard_normal((2, 40, K))).astype(np.complex64); w = o.hamming(nfft)
    y = nx.istft(torch.from_numpy(z1).cuda(), torch.from_numpy(w).cuda(), overlap_length=nfft - hop, fft_length=nfft, onesided=True)
    zz = z1.copy(); zz[..., 0] = zz[..., 0].real; zz[..., -1] = zz[..., -1].real
    full = np.concatenate([zz, np.conj(zz[..., -2:0:-1])], axis=-1)
    yo = o.istft_fast(full, w, overlap_length=nfft - hop, fft_length=nfft); chk(f"istft_c2r {nfft}/{hop}", y, yo.real.astype(np.float32))
# post-ops: median (shared-core, network, rank-counting kernels), wiener, argrel
t = rng.standard_normal((40, 70)).astype(np.float32)
for ks in [(1, 17), (9, 1), (1, 4), (3, 3), (5, 5), (7, 10), (1, 66)]:
    got = nx.Filters.median(torch.from_numpy(t).cuda(), ks).cpu().numpy(); assert np.array_equal(got, o.median(t, ks)), ks
print("median OK")
for ks, nz in [((3, 3), None), ((5, 2), 0.3)]:
    chk(f"wiener {ks}", nx.Filters.wiener(torch.from_numpy(t).cuda(), kernel_size=ks, noise=nz), o.wiener(t, ks, nz), tol=1e-6)
ti = rng.integers(-3, 4, size=(33, 130)).astype(np.float32)
for axis, order in [(0, 1), (1, 3)]:
    r = nx.PeakFinding.argrelmax(torch.from_numpy(ti).cuda(), axis=axis, order=order)
    idx, valid = o.argrelmax(ti, axis=axis, order=order)
    assert int(r["valid_indices"].cpu()) == valid and np.array_equal(r["indices"].cpu().numpy(), idx)
print("argrel OK")
# standalone framing ops (32-bit / 128-bit kernels) and the log-mel host entry's chunk pipeline
xf = rng.standard_normal((3, 4000)).astype(np.float32)
for N, st_, pad in [(64, 16, "valid"), (64, 16, "reflect"), (64, 12, "same"), (8, 4, [(8, 12)])]:
    assert np.array_equal(nx.as_windowed(torch.from_numpy(xf).cuda(), window_length=N, stride=st_, padding=pad).cpu().numpy(), o.as_windowed(xf, N, st_, pad))
tf = rng.standard_normal((2, 50, 128)).astype(np.float32)
for ov in (0, 64, 96, 127):
    chk(f"overlap_and_add {ov}", nx.overlap_and_add(torch.from_numpy(tf).cuda(), overlap_length=ov), o.overlap_and_add(tf, ov))
tc = (tf + 1j * rng.standard_normal(tf.shape)).astype(np.complex64)
chk("overlap_and_add c64", torch.view_as_real(nx.overlap_and_add(torch.from_numpy(tc).cuda(), overlap_length=96)), np.stack([o.overlap_and_add(tc, 96).real, o.overlap_and_add(tc, 96).imag], -1))
print("framing OK")
xm = rng.standard_normal((11, 20 * 1024)).astype(np.float32); wm = o.hann(1024)
kwm = dict(overlap_length=768, fft_length=1024, sampling_rate=16000)
mh = nx.stft_mel(xm, wm, mel_bins=80, **kwm)
zo, _, _ = o.stft_fast(xm, wm, **kwm)
chk("stft_mel host entry (11 channels -> 8 chunks)", mh, np.stack([o.stft_to_mel(zo[c], 16000, 1024, 80) for c in range(11)]))
# round 2: real-packed FIR kernel with an even tap count (odd K - 1: scalar stores), an unaligned row, a partitioned filter
for K, L in [(2048, 30001), (1000, 25000), (9001, 40000)]:
    x = rng.standard_normal((3, L)).astype(np.float32); taps = (rng.standard_normal(K) / np.sqrt(K)).astype(np.float32)
    for mode in ("full", "valid"):
        y = nx.convolution.convolve(torch.from_numpy(x).cuda(), torch.from_numpy(taps).cuda()[None, :], mode=mode, method="fft")
        full = oaconvolve(x.astype(np.float64), taps.astype(np.float64)[None, :], mode=mode, axes=-1)
        chk(f"fir r2c K={K} {mode}", y, full.astype(np.float32))
# complex-data STFT (plane split + combine), host pipelines on pageable numpy buffers (pinned rings, unstage + mirror)
xc = (rng.standard_normal((3, 9000)) + 1j * rng.standard_normal((3, 9000))).astype(np.complex64); wc = o.hann(512)
kwc = dict(overlap_length=384, fft_length=512, sampling_rate=48000)
zc, _, _ = nx.stft(torch.from_numpy(xc).cuda(), torch.from_numpy(wc).cuda(), **kwc)
zco, _, _ = o.stft_fast(xc, wc, **kwc); chk("stft complex data", torch.view_as_real(zc), np.stack([zco.real, zco.imag], -1))
xh = rng.standard_normal((5, 300_000)).astype(np.float32); wh = o.hann(1024)
kwh = dict(overlap_length=768, fft_length=1024, sampling_rate=48000)
zh, _, _ = nx.stft(xh, wh, **kwh)
zd, _, _ = nx.stft(torch.from_numpy(xh).cuda(), torch.from_numpy(wh).cuda(), **kwh)
assert np.array_equal(zh.view(np.float32), torch.view_as_real(zd).cpu().numpy().reshape(zh.shape[0], zh.shape[1], -1)); print("stft host (pageable) == device OK")
yh = nx.istft(zh, wh, **kwh); yd = nx.istft(zd, torch.from_numpy(wh).cuda(), **kwh)
assert np.array_equal(yh.view(np.float32), torch.view_as_real(yd).cpu().numpy().reshape(yh.shape[0], -1)); print("istft host (pageable) == device OK")
th = (rng.standard_normal(2049) / 45).astype(np.float32)
fh = nx.convolution.convolve(xh, th[None, :], mode="same", method="fft")
fd = nx.convolution.convolve(torch.from_numpy(xh).cuda(), torch.from_numpy(th).cuda()[None, :], mode="same", method="fft")
assert np.array_equal(fh, fd.cpu().numpy()); print("fir host (pageable) == device OK")
torch.cuda.synchronize(); print("all cases done")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python /tmp/san_cases.py > $OUT/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $OUT/memcheck.log; grep -c "OK" $OUT/memcheck.log; grep -E "ERROR SUMMARY|Invalid|FAIL" $OUT/memcheck.log | head -5
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 1 python /tmp/san_cases.py > $OUT/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $OUT/racecheck.log; grep -E "RACECHECK SUMMARY|hazard|FAIL" $OUT/racecheck.log | sort | uniq -c | head -10
