"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports
every symbol include/ced_abi.h declares, the host C library exports the
reference's 27 symbols, host-only helpers behave like the
reference's, and nothing silently falls back to the CPU."""
import os
import re
import subprocess

import numpy as np
import pytest

import convolutionalencdec_b200 as ced
from conftest import ROOT

REFERENCE_SYMBOLS = """g resetConvEncoder initConvEncoder convEnc computeEncOutputSegment convEncOneInput
bitReverseGenerator viterbiDecoderHard swapViterbiArrays viterbiConfigCheck viterbiInit resetViterbiDecoderHard
calcHammingDist argminPathMetrics argminNodeMetrics argmin2 argmin4 argmin8 argmin16 argmin32 argmin64
viterbiDecoderHardButterflyk1 viterbiInitButterflyk1 resetViterbiDecoderHardButterflyk1 minMetricGeneric
unpackBigToLittleEndian unpackLittleToLittleEndian""".split()


def _has_gpu():
    return ced.load_abi().ced_device_count() > 0


def test_abi_exports_every_declared_symbol():
    lib = ced.load_abi()
    declared = ced.exported_abi_symbols()
    assert len(declared) >= 19
    out = subprocess.run(["nm", "-D", "--defined-only", ced.lib_path()], capture_output=True, text=True, check=True)
    exported = {line.split()[-1] for line in out.stdout.splitlines() if line.strip()}
    for name in declared:
        assert name in exported, name
        assert hasattr(lib, name)


@pytest.mark.parametrize("params", ["k7", "k3"])
def test_dropin_exports_reference_symbols(params):
    assert len(REFERENCE_SYMBOLS) == 27
    out = subprocess.run(["nm", "-D", "--defined-only", ced.lib_path("libconvencdec_%s.so" % params)],
                         capture_output=True, text=True, check=True)
    exported = {line.split()[-1] for line in out.stdout.splitlines() if line.strip()}
    missing = [s for s in REFERENCE_SYMBOLS if s not in exported]
    assert not missing, missing
    # the extension header of the same library (include/viterbiDecoderQueue.h)
    with open(os.path.join(ROOT, "include", "viterbiDecoderQueue.h")) as f:
        text = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b((?:viterbiQueue|convEncQueue)[A-Za-z]+)\s*\(", text))
    assert len(declared) == 8 and declared <= exported, declared - exported


def test_host_tables_match_reference_values(golden):
    api = ced.RefApi("k7")
    assert (api.K, api.n, api.N, api.g) == (7, 2, 64, [0o113, 0o171])
    enc = api.encoder()
    enc.resetConvEncoder()
    enc.initConvEncoder()
    assert enc.polynomials().tolist() == golden["polys"].tolist() == [0x69, 0x4F]
    dec = api.decoder()
    dec.VITERBI_RESET()          # on "uninitialised" memory, before INIT (speedDecode.c:63-65)
    dec.VITERBI_INIT()
    assert api.viterbiConfigCheck() == 0
    assert dec.edgeCodedBitsSymm().tolist() == golden["edge_symm"].tolist()
    assert dec.nodeMetricsCur().tolist() == [0] + [65] * 63
    assert api.calcHammingDist(0b1011, 0b0001, 2) == 1 and api.calcHammingDist(0xFF, 0x0F, 8) == 4


def test_k3_params_accept_nonsymmetric_generators():
    api = ced.RefApi("k3")
    assert (api.K, api.N, api.g) == (3, 4, [0b111, 0b110])
    assert api.viterbiConfigCheck() == 0   # the reference exit(1)s here
    dec = api.decoder()
    dec.VITERBI_RESET()
    assert dec.nodeMetricsCur().tolist() == [0, 5, 5, 5]   # handTraced.c:72-75


def test_no_cpu_fallback_without_gpu():
    if _has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(ced.CedError):
        ced.Context(0)
    # the per-frame API follows the reference's error convention: message + exit(1)
    code = ("import numpy as np, convolutionalencdec_b200 as ced\n"
            "e = ced.RefApi('k7').encoder(); e.resetConvEncoder(); e.initConvEncoder()\n"
            "e.convEnc(np.zeros(4, dtype=np.uint8), True)\nprint('UNREACHABLE')\n")
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 1 and "UNREACHABLE" not in r.stdout and "GPU encoder failed" in r.stdout


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "convolutionalencdec_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".c", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"\boracle\b|orc_|refh_|swar_sim", text), os.path.join(dirpath, f)
    for lib in ("libced_cuda.so", "libconvencdec_k7.so", "libconvencdec_k3.so"):
        out = subprocess.run(["ldd", ced.lib_path(lib)], capture_output=True, text=True).stdout
        assert "oracle" not in out


def test_reference_drivers_link_unchanged():
    bindir = os.path.join(ROOT, "drivers", "_bin")
    if not os.path.isdir(bindir):
        pytest.skip("drivers not built (needs /root/reference at build time)")
    for name in ("handTraced", "berTestK7", "speedDecode", "speedEncode"):
        path = os.path.join(bindir, name)
        assert os.path.exists(path), name
        out = subprocess.run(["nm", "-D", "--undefined-only", path], capture_output=True, text=True).stdout
        assert "convEnc" in out  # resolved from the drop-in library, not compiled in


def test_sharding_covers_every_frame_once():
    from convolutionalencdec_b200.sharding import shard_range
    for n in (0, 1, 7, 65536, 2 ** 22 + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_host_side_symbol_packing_matches_the_format_definition():
    """ced_host_pack_symbols (the transfer compression inside ced_decode_batch_host): segment t -> bits
    2*(t%4).. of byte t/4, low two bits of each byte only; any length, stride and thread count."""
    import ctypes as C
    lib = ced.load_abi()
    rng = np.random.default_rng(12)
    for segs, frames, pad, threads in ((1, 3, 0, 1), (7, 5, 2, 2), (31, 9, 1, 3), (32, 4, 0, 4), (70, 33, 5, 8),
                                       (4102, 64, 10, 16), (4102, 7, 0, 5), (262, 1000, 0, 7)):
        raw = rng.integers(0, 256, (frames, segs + pad), dtype=np.uint8)
        pb = (segs + 3) // 4
        out = np.full((frames, pb + 3), 0xEE, dtype=np.uint8)
        rc = lib.ced_host_pack_symbols(raw.ctypes.data, raw.strides[0], frames, segs, out.ctypes.data, out.strides[0],
                                       threads)
        assert rc == 0
        padded = np.zeros((frames, pb * 4), dtype=np.uint8)
        padded[:, :segs] = raw[:, :segs] & 3
        q = padded.reshape(frames, pb, 4)
        want = q[..., 0] | (q[..., 1] << 2) | (q[..., 2] << 4) | (q[..., 3] << 6)
        assert np.array_equal(out[:, :pb], want), (segs, frames)
        assert (out[:, pb:] == 0xEE).all()
