import numpy as np


def synth(shape, seed, fs=48000.0):
    """BASELINE.md inputs: 0.25*N(0,1) + 0.5 sin(2 pi 440 t) + 0.5 sin(2 pi 3000 t), f32."""
    rng = np.random.default_rng(seed)
    L = shape[-1]
    t = np.arange(L, dtype=np.float64) / fs
    tone = 0.5 * np.sin(2 * np.pi * 440.0 * t) + 0.5 * np.sin(2 * np.pi * 3000.0 * t)
    x = 0.25 * rng.standard_normal(shape) + tone
    return x.astype(np.float32)


def frame_rel_err(got, want):
    """per-frame max|got - want| / max|want| (the parity metric of BASELINE.md / SURVEY 8d)."""
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    if got.size == 0:
        return 0.0
    num = np.abs(got.astype(np.complex128) - want.astype(np.complex128)).max(axis=-1)
    den = np.abs(want.astype(np.complex128)).max(axis=-1)
    den = np.where(den > 0, den, 1.0)
    return float((num / den).max())


TOL = 1.0e-5  # north_star: outputs match within 1e-5 relative fp32
