import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    # Built artefacts are git-ignored: build whatever is missing (nvcc cross-compiles without a GPU).
    import shutil
    import subprocess
    needed = {"cuda": "convolutionalencdec_b200/libced_cuda.so", "host": "convolutionalencdec_b200/libconvencdec_k7.so",
              "oracle": "oracle/libced_oracle.so", "hostsim": "tests/hostsim/libswar_sim.so"}
    missing = [t for t, path in needed.items() if not os.path.exists(os.path.join(ROOT, path))]
    if missing and shutil.which("make") and (shutil.which("nvcc") or "cuda" not in missing):
        subprocess.run(["make", "-C", ROOT] + missing, check=False, stdout=subprocess.DEVNULL)


@pytest.fixture(autouse=True)
def _pin_the_batch_kernels(request, monkeypatch):
    """ced_decode_batch hands small batches (K = 6, 7; up to ~3 rounds of one-warp CTAs) to the warp-per-frame kernel
    (csrc/warp_frame.cu).  That kernel has its own module (test_gpu_small_batch.py) and its share of the randomised soak
    (tools/fuzz_parity.py flips the switch per case); every other module names the thread-per-frame kernels it checks
    -- mostly on batches of a few dozen frames, so that the oracle finishes in seconds -- and keeps them selected."""
    if request.module.__name__.split(".")[-1] != "test_gpu_small_batch":
        monkeypatch.setenv("CED_WARP_FRAME_MAX", "0")
    yield


def bsc(rng, segs, p, junk_upper_bits=False):
    """Flip each of the 2 coded bits of every byte-per-segment symbol with probability p."""
    flips = rng.random(segs.shape + (2,)) < p
    out = segs ^ (flips[..., 0].astype(np.uint8) | (flips[..., 1].astype(np.uint8) << 1))
    if junk_upper_bits:
        out = out | (rng.integers(0, 64, segs.shape, dtype=np.uint8) << 2)
    return out


@pytest.fixture(scope="session")
def port():
    import oracle
    return oracle.port()


@pytest.fixture(scope="session")
def ref():
    import oracle
    r = oracle.ref()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this machine)")
    return r


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "k7_reference_vectors.npz")
    return np.load(path)
