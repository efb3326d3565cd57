import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _ensure_library():
    """The C-ABI library is a build artefact (git-ignored): build it when it is missing or was not built from
    the sources in the tree (nxs_build_info's source hash), so that a fresh checkout tests what it contains."""
    import shutil
    import subprocess

    lib = os.path.join(ROOT, "nx_signal_b200", "lib", "libnxsignal_b200.so")
    stale = not os.path.exists(lib)
    if not stale:  # asked in a child process, so that this one never maps a library that is about to be replaced
        r = subprocess.run([sys.executable, "-c", "from nx_signal_b200 import _lib; print(_lib.build_info()['lib_matches_tree'])"],
                           cwd=ROOT, capture_output=True, text=True)
        stale = r.returncode != 0 or r.stdout.strip() == "False"
    if stale and (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")):
        subprocess.run(["make", "-C", os.path.join(ROOT, "nx_signal_b200", "csrc"), "-j8"], check=True,
                       stdout=subprocess.DEVNULL)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    if not hasattr(config, "workerinput"):  # once, not in every xdist worker
        _ensure_library()


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
