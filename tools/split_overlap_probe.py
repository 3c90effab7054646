"""Experiment: does splitting ONE batch over k contexts/streams (forward of one part overlapping the traceback of
another, joined at the end of every batch) beat a single decode?  Prints Gbit/s per split pattern."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

frames, bits = 1 << 16, 4096
T = bits + 6
code = ced.K7_DEFAULT
main = ced.Context(0)
msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
main.random_bytes(msgs, seed=1)
segs = main.encode_batch(code, msgs, seg_stride=(T + 15) // 16 * 16)
main.bsc_channel(segs, T, 2, 0.0377, seed=2)
out = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
torch.cuda.synchronize()
ctxs = [ced.Context(0) for _ in range(4)]
streams = [torch.cuda.Stream() for _ in range(4)]
base = torch.cuda.Stream()
for pattern in ([1.0], [0.5, 0.5], [0.34, 0.66], [0.25, 0.75], [0.25, 0.25, 0.5], [0.2, 0.3, 0.5], [0.25] * 4,
                [0.15, 0.25, 0.6]):
    edges, acc = [0], 0.0
    for f in pattern:
        acc += f
        edges.append(min(frames, int(round(acc * frames / 64)) * 64))
    edges[-1] = frames
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 12
    for rep in range(reps + 3):
        if rep == 3:
            e0.record(base)
        fork = torch.cuda.Event()
        fork.record(base)
        for i in range(len(pattern)):
            streams[i].wait_event(fork)
            a, b = edges[i], edges[i + 1]
            ctxs[i].decode_batch(code, segs[a:b], bits, out=out[a:b], stream=streams[i])
            j = torch.cuda.Event()
            j.record(streams[i])
            base.wait_event(j)
    e1.record(base)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("split %-24s %.3f ms  %.1f Gbit/s" % (pattern, ms, frames * bits / ms / 1e6))
assert torch.equal(out, main.decode_batch(code, segs, bits))
