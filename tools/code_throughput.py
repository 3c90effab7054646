"""Decode throughput per code: the two compile-time K=7 codes, run-time K=7 generators through the
step-table SWAR kernel, and other K / n through the table-driven SWAR kernels of swar_generic.cu (CED_SWAR_GENERIC=0:
the one-warp-per-frame kernel they replaced)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

ctx = ced.Context(0)
bits = 4096
for K, g, frames in ((7, (0o113, 0o171), 1 << 16), (7, (0o133, 0o171), 1 << 16), (7, (0o171, 0o133), 1 << 16),
                     (7, (0o117, 0o155), 1 << 16), (7, (0o133, 0o171, 0o165), 1 << 16),
                     (7, (0o133, 0o170), 1 << 16), (7, (0o133, 0o145, 0o174), 1 << 16),
                     (3, (7, 6), 1 << 16), (3, (7, 5, 3), 1 << 16), (4, (0o15, 0o17), 1 << 16), (5, (0o23, 0o35), 1 << 16),
                     (5, (0o25, 0o33, 0o37), 1 << 16), (6, (0o53, 0o75), 1 << 16), (8, (0o247, 0o371), 1 << 16),
                     (9, (0o561, 0o753), 1 << 16), (9, (0o557, 0o663, 0o711), 1 << 16)):
    code = ced.Code(K, g)
    T = bits + K - 1
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=3)
    segs = ctx.encode_batch(code, msgs, seg_stride=(T + 15) // 16 * 16)
    ctx.bsc_channel(segs, T, len(g), 0.03, seed=4)
    out = torch.empty_like(msgs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        if rep == 1:
            e0.record()
        ctx.decode_batch(code, segs, bits, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    errs = int((out != msgs).sum())
    print("K=%d g=%s frames=%d: %8.1f Gbit/s (%.3f ms), %d wrong bytes" % (
        K, [oct(x) for x in g], frames, frames * bits / ms / 1e6, ms, errs))
ctx.close()

# rate-k/n codes with k > 1 (csrc/radix_k.cu, one warp per frame)
ctx = ced.Context(0)
for K, k, g, frames in ((3, 2, (0o27, 0o75, 0o72), 1 << 16), (3, 2, (0o53, 0o75), 1 << 16), (4, 2, (0o236, 0o155, 0o337), 1 << 16),
                        (5, 2, (0o1236, 0o0155, 0o1337), 1 << 15), (5, 2, (0o1236, 0o0155, 0o1337, 0o1701), 1 << 13)):
    code = ced.Code(K, g)
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=3)
    segs = ctx.encode_batch_k(code, k, msgs)
    out = torch.empty_like(msgs)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        if rep == 1:
            e0.record()
        ctx.decode_batch_k(code, k, segs, bits, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    print("K=%d k=%d g=%s frames=%d: %8.1f Gbit/s (%.3f ms), %d wrong bytes" % (
        K, k, [oct(x) for x in g], frames, frames * bits / ms / 1e6, ms, int((out != msgs).sum())))
ctx.close()
