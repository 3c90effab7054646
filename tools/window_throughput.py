"""Throughput of the continuous-stream decoder (ced_decode_window_batch) for several slice lengths and depths,
device-resident symbols, CUDA-event timing; prints the survivor scratch a call keeps in flight."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

streams = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
ctx = ced.Context(0)
code = ced.K7_DEFAULT
for slice_segs, depth in ((96 * 4, 48), (96 * 10, 48), (96 * 10, 96), (96 * 43, 48), (96 * 43, 96)):
    slices = max(2, 4128 // slice_segs)
    total = slices * slice_segs
    msgs = torch.empty((streams, total // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=3)
    segs = ctx.encode_batch(code, msgs, seg_stride=(total + 6 + 15) // 16 * 16)
    ctx.bsc_channel(segs, total + 6, 2, 0.0377, seed=4)
    dec = ctx.window_decoder(code, streams, depth)
    out = torch.empty((streams, (slice_segs + depth) // 8 + 1), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(3):
        dec.pos = 0
        if rep == 2:
            e0.record()
        for i in range(slices):
            dec.push(segs[:, i * slice_segs:(i + 1) * slice_segs], out=out)
        if rep == 2:
            e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("streams %d slice %5d depth %3d: %7.1f Gbit/s  (%.3f ms per slice, survivors in flight %.0f MB, carry %.1f MB)"
          % (streams, slice_segs, depth, streams * total / ms / 1e6, ms / slices,
             streams * (slice_segs + depth) * 8 / 1e6, dec.carry.numel() / 1e6))
ctx.close()
