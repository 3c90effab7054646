"""One ced_decode_batch_softq call over 2^19 frames x 4096 bits (waves pipelined inside the call), next to the hard decoder on
the same buffer.   python tools/softq_big.py"""
import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
import convolutionalencdec_b200 as ced
frames, bits, T = 1 << 19, 4096, 4102
ctx = ced.Context(0)
msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
ctx.random_bytes(msgs, seed=1)
syms = torch.empty((frames, 4112), dtype=torch.uint8, device="cuda")
for a in range(0, frames, 1 << 16):
    segs = ctx.encode_batch(ced.K7_DEFAULT, msgs[a:a + (1 << 16)], seg_stride=4112)
    soft = ctx.awgn_channel(segs, T, 3.0, seed=7, first_frame=a)
    ctx.quantize_soft(soft, T, 0.6 * 32.0 * 10 ** (-3 / 20), out=syms[a:a + (1 << 16)])
out = torch.empty_like(msgs)
for mode in ("softq", "hard"):
    fn = (lambda: ctx.decode_batch_softq(ced.K7_DEFAULT, syms, bits, out=out)) if mode == "softq" else \
         (lambda: ctx.decode_batch(ced.K7_DEFAULT, syms, bits, out=out))
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 4
    print(mode, "one call over 2^19 frames: %.1f Gbit/s" % (frames * bits / ms / 1e6), "wrong bytes", int((out != msgs).sum()) if mode == "softq" else "-")
