"""The drop-in boundary from plain C (no Python, no torch in the process): tests/c_abi/abi_smoke.c is what
a NIF does -- malloc'ed buffers, nxs_ctx_create, the _host entries -- compiled against include/nxsignal_b200.h
and linked to the in-tree libnxsignal_b200.so.  Without a GPU it must report NXS_ENODEVICE and nothing else
(there is no CPU path); on a B200 it checks STFT / ISTFT / FIR against naive double-precision sums."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "nx_signal_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "abi_smoke")
    r = subprocess.run(["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), "-L" + LIBDIR, "-lnxsignal_b200",
                        "-Wl,-rpath," + LIBDIR, "-lm", "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_program_links_and_reports_no_device_without_a_gpu(tmp_path):
    if _has_gpu():
        pytest.skip("a GPU is present: see test_c_program_runs_the_host_entries")
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3, (r.returncode, r.stdout, r.stderr)
    assert "src_sha256=" in r.stdout and "no CUDA device" in r.stdout


@pytest.mark.gpu
@pytest.mark.skipif(shutil.which("gcc") is None, reason="needs gcc")
def test_c_program_runs_the_host_entries(tmp_path):
    r = subprocess.run([_build(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert "c abi smoke ok" in r.stdout
