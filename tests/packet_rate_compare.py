"""TEST-SIDE MEASUREMENT SCRIPT (not collected by pytest): the reference's speedDecode loop -- 16 random
2048-bit packets, pre-encoded, one VITERBI_DECODER_HARD(last=true) call per packet (speedDecode.c:37-119)
-- timed (a) on ONE host core with the unmodified reference C (oracle/_ref, the loop in
oracle/ref_harness.c) and (b) through this repo's drop-in per-packet API on the GPU.

    python tests/packet_rate_compare.py [seconds]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import convolutionalencdec_b200 as ced  # noqa: E402
import oracle  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 2.0
R = oracle.ref()
rng = np.random.default_rng(314)
for bits in (2048, 4096):
    msgs = rng.integers(0, 256, (16, bits // 8), dtype=np.uint8)
    segs = R.encode_batch(msgs)
    T = bits + 6
    done, el = R.speed_decode(segs, T, 1, seconds)
    api = ced.RefApi("k7")
    dec = api.decoder(); dec.VITERBI_RESET(); dec.VITERBI_INIT()
    lib, p, out = api.lib, dec.p, np.zeros(bits // 8 + 8, dtype=np.uint8)
    ptrs = [segs[i].ctypes.data_as(oracle._u8p) for i in range(16)]
    outp = out.ctypes.data_as(oracle._u8p)
    for i in range(64):
        lib.viterbiDecoderHardButterflyk1(p, ptrs[i % 16], outp, T, True)
    assert np.array_equal(out[:bits // 8], msgs[63 % 16])
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        for i in range(16):
            lib.viterbiDecoderHardButterflyk1(p, ptrs[i], outp, T, True)
        n += 16
    dt = time.perf_counter() - t0
    print("%d-bit packets: reference C on 1 host core %.1f Mbit/s | drop-in API on the GPU %.1f Mbit/s (%.1f us per call)"
          % (bits, done / el / 1e6, n * bits / dt / 1e6, dt / n * 1e6))
