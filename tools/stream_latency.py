"""Per-call latency of the reference-named per-frame API (VITERBI_DECODER_HARD / convEnc through the
GPU) for several packet lengths: separates the fixed host+launch overhead from the per-step cost."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

api = ced.RefApi("k7")
enc = api.encoder(); enc.resetConvEncoder(); enc.initConvEncoder()
dec = api.decoder(); dec.VITERBI_RESET(); dec.VITERBI_INIT()
rng = np.random.default_rng(0)
for bits in (8, 256, 2048, 4096, 16384):
    msg = rng.integers(0, 256, bits // 8, dtype=np.uint8)
    segs = enc.convEnc(msg, True)
    for _ in range(20):
        out = dec.VITERBI_DECODER_HARD(segs, True, max_bytes=4096)
    assert np.array_equal(out, msg)
    n = 200
    t0 = time.perf_counter()
    for _ in range(n):
        dec.VITERBI_DECODER_HARD(segs, True, max_bytes=4096)
    dt = (time.perf_counter() - t0) / n
    t0 = time.perf_counter()
    for _ in range(n):
        enc.convEnc(msg, True)
    de = (time.perf_counter() - t0) / n
    print("bits %6d: decode %8.1f us/call (%6.2f Mbit/s)   encode %7.1f us/call (%7.1f Mbit/s)"
          % (bits, dt * 1e6, bits / dt / 1e6, de * 1e6, bits / de / 1e6))
