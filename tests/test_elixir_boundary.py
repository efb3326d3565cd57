"""The Elixir side of the drop-in boundary cannot be compiled here -- there is
no BEAM in the image -- so it is checked as far as the image allows:

* elixir/c_src/nxsignal_nif.c is syntax- and type-checked by gcc against tests/stubs/erl_nif.h (OTP's
  signatures for exactly the calls the file makes) and the real include/nxsignal_b200.h, so every
  nxs_* call in it matches the C ABI's prototypes;
* every NIF stub the Elixir module declares exists in the C file's funcs[] with the same arity, and
  the other way round;
* every `_host` export of the C ABI is reachable from a NIF;
* the shim exposes every public head SURVEY 8b lists, with the reference's names and arities.
"""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NIF_C = os.path.join(ROOT, "elixir", "c_src", "nxsignal_nif.c")
SHIM = os.path.join(ROOT, "elixir", "lib", "nx_signal_b200.ex")
HEADER = os.path.join(ROOT, "include", "nxsignal_b200.h")


def _c_funcs():
    src = open(NIF_C).read()
    table = src[src.index("static ErlNifFunc funcs[]"):]
    return {m.group(1): int(m.group(2)) for m in re.finditer(r'\{"(\w+)",\s*(\d+),\s*(\w+),', table)}


def _ex_nifs():
    src = open(SHIM).read()
    mod = src[src.index("defmodule NxSignalB200.NIF do"):]
    mod = mod[:mod.index("\nend\n")]
    out = {}
    for m in re.finditer(r"def (\w+)\(([^)]*)\)\s*,?\s*(?:do:|\n\s*do:)\s*:erlang\.nif_error", mod):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_nif_source_type_checks_against_the_c_abi():
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only",
                        "-I" + os.path.join(ROOT, "tests", "stubs"), "-I" + os.path.join(ROOT, "include"), NIF_C],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_every_elixir_nif_stub_has_a_c_entry_with_the_same_arity():
    c, ex = _c_funcs(), _ex_nifs()
    assert len(ex) >= 17
    assert ex == c, (sorted(set(ex.items()) ^ set(c.items())))


def test_every_host_export_is_reachable_from_a_nif():
    header = open(HEADER).read()
    hosts = set(re.findall(r"\b(nxs_\w+_host)\s*\(", header))
    closed_forms = {"nxs_window_f32", "nxs_firwin_f32", "nxs_fft_frequencies_f32", "nxs_mel_filters_f32",
                    "nxs_stft_times_f32", "nxs_num_frames", "nxs_fir_out_len"}
    src = open(NIF_C).read()
    missing = [s for s in sorted(hosts | closed_forms) if not re.search(r"\b" + s + r"\s*\(", src)]
    assert not missing, missing
    assert "enif_mutex_lock" in src and "ERL_NIF_DIRTY_JOB_IO_BOUND" in src


def test_nif_validates_sizes_before_allocating():
    """every NIF that inspects a binary compares its size with the declared dimensions, and every result
    allocation is preceded by an overflow-checked byte count"""
    src = open(NIF_C).read()
    bodies = re.split(r"\nstatic ERL_NIF_TERM ", src)[1:]
    for body in bodies:
        name = body.split("(")[0]
        if "enif_inspect_binary" in body:
            assert re.search(r"\.size\s*!=|\(int64_t\)\w+\.size\s*!=", body), f"{name}: binary size not checked"
        if "enif_make_new_binary" in body:
            first_alloc = body.index("enif_make_new_binary")
            assert "bytes3(" in body[:first_alloc] or "mul_ok(" in body[:first_alloc], f"{name}: unchecked allocation size"


# the reference's public heads on the accelerated path: module -> {name: arity with all defaults given}
HEADS = {
    "NxSignalB200": {"stft": 3, "istft": 3, "as_windowed": 2, "overlap_and_add": 2, "fft_frequencies": 2,
                     "mel_filters": 4, "stft_to_mel": 3},
    "NxSignalB200.Windows": {"rectangular": 2, "bartlett": 2, "triangular": 2, "blackman": 2, "hamming": 2, "hann": 2,
                             "kaiser": 2},
    "NxSignalB200.Filters": {"firwin": 3, "median": 2, "wiener": 2},
    "NxSignalB200.Convolution": {"convolve": 3, "correlate": 3, "fftconvolve": 3},
    "NxSignalB200.PeakFinding": {"argrelmin": 2, "argrelmax": 2, "argrelextrema": 3},
}


def test_shim_exposes_the_reference_heads():
    src = open(SHIM).read()
    for mod, heads in HEADS.items():
        start = src.index(f"defmodule {mod} do")
        nxt = src.find("\ndefmodule ", start + 1)
        body = src[start: nxt if nxt > 0 else len(src)]
        for name, arity in heads.items():
            m = re.search(r"\n  def " + name + r"\(([^)]*)\)", body)
            assert m, f"{mod}.{name} missing"
            args = [a for a in re.sub(r"\\\\\s*\[\]", "", m.group(1)).split(",") if a.strip()]
            assert len(args) == arity, f"{mod}.{name}/{arity}: found {len(args)} parameters"


def test_shim_keeps_the_reference_error_texts():
    src = open(SHIM).read()
    for text in ["missing sampling_rate option",
                 "invalid :scaling, expected one of :spectrum, :psd or nil, got:",
                 ":sampling_rate is mandatory if scaling is :psd",
                 "expected an integer >= 1 or a list of integers, got:",
                 "invalid padding mode specified, padding must be one of :valid, :same, or a padding configuration, got:",
                 "overlap_length must be a number less than the window size",
                 "cutoff must be a list of frequencies, got:",
                 "cutoff must be strictly between 0 and Nyquist (exclusive), got:",
                 "requires an odd number of taps, got:",
                 "kernel shape must be of the same rank as the tensor",
                 "kernel_size must be an integer or tuple",
                 "expected mode to be one of [:full, :same, :valid], got:",
                 "expected method to be one of [:direct, :fft], got:",
                 "Rank of in1 and in2 must be equal.",
                 "For :valid mode, one must be at least as large as the other in every dimension"]:
        assert text in src, text
