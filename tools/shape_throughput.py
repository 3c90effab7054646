"""Decoded Gbit/s of ced_decode_batch / ced_encode_batch for the default code over frame shapes with the same total
number of information bits (2^28): does the batch path hold its rate for short and for long frames?"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

ctx = ced.Context(0)
code = ced.K7_DEFAULT
total = 1 << 28
for bits in (64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384):
    frames = total // bits
    T = bits + 6
    stride = (T + 15) // 16 * 16
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=5)
    segs = torch.zeros((frames, stride), dtype=torch.uint8, device="cuda")
    out = torch.empty_like(msgs)
    res = []
    for what in ("enc", "dec"):
        def run():
            if what == "enc":
                ctx.encode_batch(code, msgs, out=segs)
            else:
                ctx.decode_batch(code, segs, bits, out=out)
        for _ in range(2):
            run()
        ctx.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.current_stream()
        torch.cuda.synchronize()
        n = 5
        import time
        t0 = time.perf_counter()
        for _ in range(n):
            run()
        ctx.sync()
        res.append(total * n / (time.perf_counter() - t0) / 1e9)
    assert torch.equal(out, msgs)
    print("%6d bits x %8d frames: encode %7.1f Gbit/s, decode %6.1f Gbit/s" % (bits, frames, res[0], res[1]), flush=True)
    del msgs, segs, out
ctx.close()
