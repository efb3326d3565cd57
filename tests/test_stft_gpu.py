"""GPU parity of the fused STFT kernel (through the C ABI) against the oracle.
Tolerance: per-frame max|gpu - oracle| / max|oracle| <= 1e-5 (north_star's 1e-5 relative
fp32; element-wise relative error is meaningless at near-zero bins)."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from oracle import nxsignal_oracle as o
from tests.util import TOL, frame_rel_err, synth

pytestmark = pytest.mark.gpu


def test_doctest_stft_rect2():  # lib/nx_signal.ex:46-65 (generic DFT path, nfft = 2)
    z, t, f = nx.stft(np.arange(4, dtype=np.int32), nx.windows.rectangular(2), overlap_length=1, fft_length=2,
                      sampling_rate=400)
    np.testing.assert_array_equal(z, np.array([[1, -1], [3, -1], [5, -1]], dtype=np.complex64))
    np.testing.assert_array_equal(t, np.array([0.0025, 0.005, 0.0075], dtype=np.float32))
    np.testing.assert_array_equal(f, np.array([0.0, 200.0], dtype=np.float32))


def test_doctest_stft_reflect_nfft16():  # lib/nx_signal.ex:465-471
    kw = dict(overlap_length=2, fft_length=16, sampling_rate=8.0e3, window_padding="reflect")
    z, _, _ = nx.stft(np.arange(10, dtype=np.int32), nx.windows.hann(4), **kw)
    zo, _, _ = o.stft(np.arange(10, dtype=np.int32), o.hann(4), **kw)
    assert z.shape == (6, 16)
    assert frame_rel_err(z, zo) <= TOL


def test_cfg1_vs_literal_oracle():
    """BASELINE config 1: 1 x 48000, nfft 1024, hop 256, Hann -- against the literal
    recursive-radix-2 f64 restatement of Nx.BinaryBackend."""
    x = synth((48000,), 1001)
    w = nx.windows.hann(1024)
    z, t, f = nx.stft(x, w, overlap_length=768, fft_length=1024, sampling_rate=48000)
    zo, to, fo = o.stft(x, o.hann(1024), overlap_length=768, fft_length=1024, sampling_rate=48000)
    assert z.shape == (184, 1024) and z.dtype == np.complex64
    assert frame_rel_err(z, zo) <= TOL
    np.testing.assert_array_equal(t, to)
    np.testing.assert_array_equal(f, fo)


@pytest.mark.parametrize("nfft", [64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
def test_every_pow2_plan(nfft):
    x = synth((3, 6 * nfft + 37), 7 + nfft)
    w = o.hann(nfft)
    z, _, _ = nx.stft(x, w, overlap_length=nfft - nfft // 4, fft_length=nfft, sampling_rate=48000)
    zo, _, _ = o.stft_fast(x, w, overlap_length=nfft - nfft // 4, fft_length=nfft, sampling_rate=48000)
    assert z.shape == zo.shape
    assert frame_rel_err(z, zo) <= TOL


@pytest.mark.parametrize("nfft", [1, 2, 3, 5, 10, 16, 30, 32, 100, 250, 1000])
def test_generic_dft_lengths(nfft):
    N = max(2, min(nfft, 24))
    x = synth((2, 400), 99 + nfft)
    w = o.hamming(N)
    z, _, _ = nx.stft(x, w, overlap_length=N // 2, fft_length=nfft, sampling_rate=100)
    zo, _, _ = o.stft(x, w, overlap_length=N // 2, fft_length=nfft, sampling_rate=100)
    assert frame_rel_err(z, zo) <= TOL


@pytest.mark.parametrize("padding", ["valid", "same", "reflect", [(5, 9)], [(0, 300)]])
@pytest.mark.parametrize("N,hop,nfft", [(256, 64, 256), (200, 50, 256), (300, 77, 256), (256, 255, 1024), (64, 64, 64)])
def test_padding_modes_and_lengths(padding, N, hop, nfft):
    x = synth((2, 3001), 5)
    w = o.blackman(N)
    kw = dict(overlap_length=N - hop, fft_length=nfft, sampling_rate=8000, window_padding=padding)
    z, t, f = nx.stft(x, w, **kw)
    zo, to, fo = o.stft_fast(x, w, **kw)
    assert z.shape == zo.shape
    assert frame_rel_err(z, zo) <= TOL
    np.testing.assert_array_equal(t, to)
    np.testing.assert_array_equal(f, fo)


@pytest.mark.parametrize("scaling", [None, "spectrum", "psd"])
def test_scaling(scaling):
    x = synth((4, 9000), 11)
    w = o.hann(512)
    kw = dict(overlap_length=384, fft_length=512, sampling_rate=16000, scaling=scaling)
    z, _, _ = nx.stft(x, w, **kw)
    zo, _, _ = o.stft_fast(x, w, **kw)
    assert frame_rel_err(z, zo) <= TOL


def test_batch_axes_and_int_input():
    x = (synth((2, 3, 5000), 3) * 1000).astype(np.int32)
    w = o.hann(128)
    z, _, _ = nx.stft(x, w, fft_length=128, sampling_rate=100)
    zo, _, _ = o.stft_fast(x.astype(np.float32), w, fft_length=128, sampling_rate=100)
    assert z.shape == (2, 3, zo.shape[-2], 128)
    assert frame_rel_err(z, zo) <= TOL


def test_device_entry_matches_host_entry():
    import torch

    x = synth((5, 40000), 21)
    w = nx.windows.hann(1024)
    zh, _, _ = nx.stft(x, w, overlap_length=768, sampling_rate=48000)
    xd = torch.from_numpy(x).cuda()
    zd, td, fd = nx.stft(xd, torch.from_numpy(w).cuda(), overlap_length=768, sampling_rate=48000)
    torch.cuda.synchronize()
    assert zd.is_cuda and zd.dtype == torch.complex64
    np.testing.assert_array_equal(zd.cpu().numpy(), zh)


def test_empty_and_ragged():
    w = nx.windows.hann(64)
    z, t, f = nx.stft(np.zeros(10, np.float32), w, sampling_rate=100)  # shorter than one frame
    assert z.shape == (0, 64) and t.shape == (0,) and f.shape == (64,)
    z, _, _ = nx.stft(synth((1, 64), 1), w, sampling_rate=100)  # exactly one frame
    zo, _, _ = o.stft_fast(synth((1, 64), 1), w, sampling_rate=100)
    assert z.shape == (1, 1, 64) and frame_rel_err(z, zo) <= TOL


def test_properties_at_scale():
    """Size-independent checks on a large shape (oracle would take minutes): linearity and
    Parseval per frame, Hermitian symmetry of the two-sided spectrum."""
    import torch

    C, L, N, H = 4, 2_000_000, 1024, 256
    g = torch.Generator(device="cuda").manual_seed(5)
    a = torch.randn(C, L, device="cuda", generator=g)
    b = torch.randn(C, L, device="cuda", generator=g)
    w = torch.from_numpy(nx.windows.hann(N)).cuda()
    za, _, _ = nx.stft(a, w, overlap_length=N - H, sampling_rate=48000)
    zb, _, _ = nx.stft(b, w, overlap_length=N - H, sampling_rate=48000)
    zab, _, _ = nx.stft(2.0 * a - 3.0 * b, w, overlap_length=N - H, sampling_rate=48000)
    lin = (zab - (2.0 * za - 3.0 * zb)).abs().amax(dim=-1) / zab.abs().amax(dim=-1)
    assert float(lin.max()) < 5e-6
    # Parseval: sum |X|^2 = nfft * sum (x w)^2
    frames = a.unfold(-1, N, H) * w
    e_t = (frames.double() ** 2).sum(-1) * N
    e_f = (za.abs().double() ** 2).sum(-1)
    assert float(((e_t - e_f).abs() / e_t).max()) < 1e-5
    herm = (za[..., 1:] - za[..., 1:].flip(-1).conj()).abs().amax(dim=-1) / za.abs().amax(dim=-1)
    assert float(herm.max()) < 1e-6


# ---- one-sided device form and the host entry's mirror pipeline ---------------------------------
@pytest.mark.parametrize("nfft", [64, 256, 512, 1024, 2048, 4096, 8192])
@pytest.mark.parametrize("padding", ["valid", "reflect"])
def test_onesided_is_the_lower_half_bit_for_bit(nfft, padding):
    import torch

    x = torch.from_numpy(synth((3, 9 * nfft + 13), 31 + nfft)).cuda()
    w = torch.from_numpy(o.hann(nfft)).cuda()
    kw = dict(overlap_length=nfft - nfft // 4, fft_length=nfft, sampling_rate=48000, window_padding=padding)
    z2, _, f2 = nx.stft(x, w, **kw)
    z1, _, f1 = nx.stft(x, w, onesided=True, **kw)
    assert z1.shape == z2.shape[:-1] + (nfft // 2 + 1,)
    assert torch.equal(torch.view_as_real(z1), torch.view_as_real(z2[..., : nfft // 2 + 1].contiguous()))
    assert torch.equal(f1, f2[: nfft // 2 + 1])
    # and the dropped half is the exact conjugate mirror
    up = torch.view_as_real(z2[..., nfft // 2 + 1:].contiguous())
    mir = torch.view_as_real(torch.conj(z2[..., 1: nfft // 2].flip(-1)).resolve_conj().contiguous())
    assert torch.equal(up, mir)


def test_onesided_generic_dft_length():
    import torch

    x = torch.from_numpy(synth((2, 300), 5)).cuda()
    w = torch.from_numpy(o.hamming(20)).cuda()
    kw = dict(overlap_length=10, fft_length=30, sampling_rate=100)
    z2, _, _ = nx.stft(x, w, **kw)
    z1, _, _ = nx.stft(x, w, onesided=True, **kw)
    assert torch.equal(torch.view_as_real(z1), torch.view_as_real(z2[..., :16].contiguous()))


@pytest.mark.parametrize("nfft,hop,C,L,padding,scaling", [
    (1024, 256, 2, 1_300_000, "valid", None),      # > 1 D2H slab per channel chunk
    (1024, 256, 3, 50_000, "reflect", "spectrum"),
    (64, 16, 5, 4_001, "same", "psd"),
    (2048, 512, 2, 70_000, "valid", None),
    (4096, 1024, 2, 70_000, "valid", None),
    (512, 200, 1, 30_001, "valid", None),          # hop not a multiple of 4: general r2c kernel
    (30, 10, 2, 500, "valid", None),               # generic DFT: two-sided on the wire
])
def test_host_entry_is_bit_identical_to_device_entry(nfft, hop, C, L, padding, scaling):
    """nxs_stft_f32_host moves bins 0..nfft/2 over PCIe and mirrors on the host; the result must
    equal the two-sided device result bit for bit (pageable numpy buffers here)."""
    import torch

    x = synth((C, L), 77 + nfft)
    w = o.hann(min(nfft, 1024)) if nfft != 30 else o.hamming(20)
    N = len(w)
    kw = dict(overlap_length=N - hop, fft_length=nfft, sampling_rate=48000, window_padding=padding, scaling=scaling)
    zh, th, fh = nx.stft(x, w, **kw)
    zd, td, fd = nx.stft(torch.from_numpy(x).cuda(), torch.from_numpy(w).cuda(), **kw)
    assert zh.shape == tuple(zd.shape)
    np.testing.assert_array_equal(zh.view(np.float32), torch.view_as_real(zd).cpu().numpy().reshape(zh.shape[:-1] + (-1,)))
    np.testing.assert_array_equal(th, td.cpu().numpy())
    zo, _, _ = o.stft_fast(x, w, **kw)
    assert frame_rel_err(zh, zo) <= TOL


def test_host_entry_pinned_buffers_and_repeat():
    """Pinned caller buffers (what bench.py's e2e leg uses), called twice on one context."""
    import torch

    from nx_signal_b200 import _arrays as A
    from nx_signal_b200 import _lib

    C_, L, nfft, hop = 4, 600_000, 1024, 256
    M = (L - nfft) // hop + 1
    xh = torch.from_numpy(synth((C_, L), 404)).pin_memory()
    w = o.hann(nfft)
    zh = torch.empty((C_, M, nfft), dtype=torch.complex64).pin_memory()
    ctx = _lib.context(0)
    for _ in range(2):
        zh.zero_()
        _lib.check(_lib.lib().nxs_stft_f32_host(ctx, A.ptr(xh), C_, L, L, w.ctypes.data, nfft, hop, nfft,
                                                _lib.PAD_VALID, 0, 0, _lib.SCALE_NONE, 48000.0, A.ptr(zh)), ctx)
        zd, _, _ = nx.stft(xh.cuda(), torch.from_numpy(w).cuda(), overlap_length=nfft - hop, fft_length=nfft,
                           sampling_rate=48000)
        assert torch.equal(torch.view_as_real(zh), torch.view_as_real(zd).cpu())


@pytest.mark.parametrize("nfft,onesided", [(4096, False), (4096, True)])
def test_large_calls_walked_in_channel_blocks_give_the_same_bits(nfft, onesided, monkeypatch):
    """launch_stft walks nfft-4096 inputs of tens of GB in channel blocks (csrc/nxs_stft.cu); NXS_STFT_SPLIT_BYTES lowers
    the block size so that a small call takes that route: same bits as the single launch, uneven last block included"""
    import torch

    Cn, L = 7, 60 * nfft + 16
    x = torch.from_numpy(synth((Cn, L), 900 + nfft)).cuda()
    w = torch.from_numpy(o.hann(nfft)).cuda()
    kw = dict(overlap_length=nfft - nfft // 4, fft_length=nfft, sampling_rate=48000)
    if onesided:
        kw["onesided"] = True
    z0 = nx.stft(x, w, **kw)[0]
    monkeypatch.setenv("NXS_STFT_SPLIT_BYTES", str(2.5 * (4 * L + 8 * ((L - nfft) // (nfft // 4) + 1) * nfft)))  # blocks of 2-3 channels
    z1 = nx.stft(x, w, **kw)[0]
    assert torch.equal(torch.view_as_real(z0), torch.view_as_real(z1))
