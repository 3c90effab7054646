"""Decode 2^16 (or argv[4]) frames x 4096 bits of ONE code three times -- a target for ncu captures of the kernels behind
other code parameters.   python tools/one_code.py K g0,g1[,g2] [k] [frames]      (generators in octal)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

K = int(sys.argv[1])
g = tuple(int(x, 8) for x in sys.argv[2].split(","))
k = int(sys.argv[3]) if len(sys.argv) > 3 else 1
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 16
bits = 4096
ctx = ced.Context(0)
code = ced.Code(K, g)
msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
ctx.random_bytes(msgs, seed=3)
if k == 1:
    T = bits + K - 1
    segs = ctx.encode_batch(code, msgs, seg_stride=(T + 15) // 16 * 16)
    ctx.bsc_channel(segs, T, len(g), 0.03, seed=4)
else:
    segs = ctx.encode_batch_k(code, k, msgs)
out = torch.empty_like(msgs)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for rep in range(3):
    if rep == 1:
        e0.record()
    if k == 1:
        ctx.decode_batch(code, segs, bits, out=out)
    else:
        ctx.decode_batch_k(code, k, segs, bits, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print("K=%d k=%d g=%s frames=%d: %.1f Gbit/s (%.3f ms)" % (K, k, [oct(x) for x in g], frames, frames * bits / ms / 1e6, ms))
ctx.close()
