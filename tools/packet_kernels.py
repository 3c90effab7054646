"""One-shot VITERBI_DECODER_HARD calls of a few packet lengths, for an ncu launch list of the
frame-parallel kernels (fpBlockKernel / fpSelectKernel):
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/packet_kernels.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

api = ced.RefApi("k7")
enc = api.encoder(); enc.resetConvEncoder(); enc.initConvEncoder()
dec = api.decoder(); dec.VITERBI_RESET(); dec.VITERBI_INIT()
rng = np.random.default_rng(0)
for bits in (2048, 4096, 16384):
    msg = rng.integers(0, 256, bits // 8, dtype=np.uint8)
    segs = enc.convEnc(msg, True)
    for _ in range(3):
        out = dec.VITERBI_DECODER_HARD(segs, True, max_bytes=4096)
    assert np.array_equal(out, msg)
print("ok")
