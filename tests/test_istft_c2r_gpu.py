"""GPU parity of the c2r ISTFT (opt-in one-sided form; C ABI
nxs_istft_c2r_f32_{dev,host}) against the oracle: y = Re(istft(ext(z))) with ext the
conjugate-mirror extension of the one-sided spectrum (lib/nx_signal.ex:582-638 on ext(z)).
Tolerance: max|gpu - oracle| / max|oracle| <= 1e-5 per channel (north_star)."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
from tests.util import TOL, synth

pytestmark = pytest.mark.gpu


def ext(z1, nfft):
    """bins 0 .. nfft/2 -> the two-sided Hermitian spectrum (DC / Nyquist imaginary parts dropped)."""
    z1 = np.asarray(z1).astype(np.complex64).copy()
    z1[..., 0] = z1[..., 0].real
    z1[..., nfft // 2] = z1[..., nfft // 2].real
    return np.concatenate([z1[..., :nfft // 2 + 1], np.conj(z1[..., nfft // 2 - 1:0:-1])], axis=-1)


def rel(got, want):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (got.shape, want.shape)
    den = np.abs(want).max(axis=-1)
    return float((np.abs(got - want).max(axis=-1) / np.where(den > 0, den, 1.0)).max())


def onesided_random(rng, shape_cm, nfft, z_ld=None):
    K = nfft // 2 + 1
    z = (rng.standard_normal(shape_cm + (K,)) + 1j * rng.standard_normal(shape_cm + (K,))).astype(np.complex64)
    if z_ld is None:
        return z, z
    buf = np.zeros(shape_cm + (z_ld,), dtype=np.complex64)
    buf[..., :K] = z
    buf[..., K:] = 7.0  # row slack must never be read as data
    return z, buf


@pytest.mark.parametrize("nfft,hopdiv", [(512, 2), (512, 4), (512, 8), (1024, 2), (1024, 4), (1024, 8), (2048, 2),
                                         (2048, 4), (2048, 8), (4096, 2), (4096, 4), (4096, 8)])
@pytest.mark.parametrize("scalar", [False, True])
def test_fast_plans_vs_oracle(nfft, hopdiv, scalar, monkeypatch):
    """default: FFT engine on packed fp32x2 arithmetic (Plan::PK); NXS_ISTFT_SCALAR=1: the scalar plans"""
    if scalar:
        monkeypatch.setenv("NXS_ISTFT_SCALAR", "1")
    rng = np.random.default_rng(nfft + hopdiv)
    hop = nfft // hopdiv
    M = 41
    z, _ = onesided_random(rng, (3, M), nfft)  # odd row length: every other row starts 8 bytes off 16
    w = o.hamming(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft)
    y = nx.istft(z, w, onesided=True, **kw)
    assert y.dtype == np.float32 and y.shape == (3, M * hop + nfft - hop)
    yo = o.istft_fast(ext(z, nfft), w, **kw)
    assert rel(y, yo.real) <= TOL
    assert np.abs(yo.imag).max() <= 1e-5 * np.abs(yo.real).max()  # ext(z) is Hermitian: the reference's imag is noise


@pytest.mark.parametrize("nfft,hop", [(256, 64), (1024, 250), (1024, 1024), (64, 16), (12, 5), (8192, 2048), (1024, 192)])
def test_other_shapes_take_the_extension_path(nfft, hop):
    rng = np.random.default_rng(nfft * 7 + hop)
    M = 23
    z, _ = onesided_random(rng, (2, M), nfft)
    w = (o.hann(nfft) + np.float32(0.1)).astype(np.float32)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft)
    y = nx.istft(z, w, onesided=True, **kw)
    yo = (o.istft_fast if nfft & (nfft - 1) == 0 else o.istft)(ext(z, nfft), w, **kw)
    assert y.shape == yo.shape
    assert rel(y, yo.real) <= TOL


def test_cuda_tensor_row_stride_and_slack():
    import torch

    rng = np.random.default_rng(11)
    nfft, hop, M = 1024, 256, 300
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft)
    for z_ld in (513, 514, 516, 600):
        z, buf = onesided_random(rng, (2, M), nfft, z_ld)
        zd = torch.from_numpy(buf).cuda()
        y = nx.istft(zd, torch.from_numpy(w).cuda(), onesided=True, **kw)
        torch.cuda.synchronize()
        assert y.dtype == torch.float32 and y.is_cuda
        yo = o.istft_fast(ext(z, nfft), w, **kw)
        assert rel(y.cpu().numpy(), yo.real) <= TOL, z_ld


def test_dc_and_nyquist_imaginary_parts_are_ignored():
    rng = np.random.default_rng(12)
    nfft, hop = 1024, 256
    z, _ = onesided_random(rng, (1, 50), nfft)
    z2 = z.copy()
    z2[..., 0] = z2[..., 0].real
    z2[..., nfft // 2] = z2[..., nfft // 2].real
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft, onesided=True)
    assert np.array_equal(nx.istft(z, w, **kw), nx.istft(z2, w, **kw))


@pytest.mark.parametrize("scaling", [None, "spectrum", "psd"])
def test_round_trip_of_the_onesided_stft(scaling):
    """cfg5 shape at reduced length.  (i) fed the oracle's spectrum (lower half), the c2r path equals the
    real part of the reference's c64 result everywhere; (ii) stft(onesided) -> istft(onesided) on the
    device reproduces x away from the edges and equals our own c64 round trip everywhere (at the
    ill-conditioned edges the result depends on the spectrum's last bits, so (ii) is not compared
    with the oracle's round trip there -- tests/test_istft_gpu.py does the same)."""
    import torch

    x = synth((3, 120_000), 1005)
    w = o.hann(1024)
    kw = dict(overlap_length=768, fft_length=1024, sampling_rate=48000, scaling=scaling)
    zo, _, _ = o.stft_fast(x, w, **kw)
    yo = o.istft_fast(zo, w, **kw)
    assert rel(nx.istft(zo[..., :513], w, onesided=True, **kw), yo.real) <= TOL
    xd, wd = torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda()
    z1, _, _ = nx.stft(xd, wd, onesided=True, **kw)
    assert z1.shape[-1] == 513
    y = nx.istft(z1, wd, onesided=True, **kw).cpu().numpy()
    n = y.shape[-1]
    assert np.abs(y[:, 1024:n - 1024] - x[:, 1024:n - 1024]).max() <= 1e-5 * np.abs(x).max()
    assert rel(y[:, 1024:n - 1024], yo.real[:, 1024:n - 1024]) <= TOL
    z2, _, _ = nx.stft(xd, wd, **kw)
    y2 = nx.istft(z2, wd, **kw).cpu().numpy()
    assert rel(y, y2.real) <= TOL


def test_long_channel_many_segments():
    rng = np.random.default_rng(13)
    nfft, hop, M = 512, 128, 6000
    z, _ = onesided_random(rng, (2, M), nfft)
    w = o.hann(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft)
    y = nx.istft(z, w, onesided=True, **kw)
    yo = o.istft_fast(ext(z, nfft), w, **kw)
    assert rel(y, yo.real) <= TOL


def test_edge_samples_meet_plain_tolerance():
    """The first / last ~0.1 N samples divide by a vanishing window energy: the f64 fix-up kernel
    (Hermitian form) keeps them inside the plain bound, element-wise relative to the channel maximum."""
    x = synth((2, 40_000), 77)
    w = o.hann(1024)
    kw = dict(overlap_length=768, fft_length=1024)
    zo, _, _ = o.stft_fast(x, w, **kw)
    yo = o.istft_fast(zo, w, **kw).real
    y = nx.istft(zo[..., :513], w, onesided=True, **kw)
    assert np.abs(y[:, :128] - yo[:, :128]).max() <= TOL * np.abs(yo).max()
    assert np.abs(y[:, -128:] - yo[:, -128:]).max() <= TOL * np.abs(yo).max()


def test_argument_errors():
    w = o.hann(1024)
    z = np.zeros((1, 4, 512), dtype=np.complex64)  # one bin short
    with pytest.raises(nx.NxSignalArgumentError):
        nx.istft(z, w, onesided=True, overlap_length=768, fft_length=1024)


@pytest.mark.parametrize("M", [1, 2, 3, 5, 9])
@pytest.mark.parametrize("nfft,hopdiv", [(1024, 4), (512, 2), (2048, 8)])
def test_very_few_frames(nfft, hopdiv, M):
    """Fewer frames than cover an interior sample: every output is an edge sample (exact normaliser path)."""
    rng = np.random.default_rng(100 * M + hopdiv)
    hop = nfft // hopdiv
    z, _ = onesided_random(rng, (2, M), nfft)
    w = o.hamming(nfft)
    kw = dict(overlap_length=nfft - hop, fft_length=nfft)
    y = nx.istft(z, w, onesided=True, **kw)
    yo = o.istft_fast(ext(z, nfft), w, **kw)
    assert y.shape == yo.shape == (2, M * hop + nfft - hop)
    assert rel(y, yo.real) <= TOL
