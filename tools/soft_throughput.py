"""Soft-decision decode throughput at the BASELINE config-2 shape (2^16 frames x 4096 bits), kernel times from
CUDA events on the launching stream (ced_ctx_set_profiling).  python tools/soft_throughput.py [frames]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import convolutionalencdec_b200 as ced

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
bits, T = 4096, 4102
ctx = ced.Context(0)
msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
ctx.random_bytes(msgs, seed=314)
segs = torch.zeros((frames, 4112), dtype=torch.uint8, device="cuda")
ctx.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
soft = ctx.awgn_channel(segs, T, 3.0, seed=2718)
out = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.decode_batch_soft(ced.K7_DEFAULT, soft, bits, out=out)
ctx.sync()
ctx.set_profiling(True)
fwd, tb = [], []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10
e0.record()
for _ in range(n):
    ctx.decode_batch_soft(ced.K7_DEFAULT, soft, bits, out=out)
    f, t = ctx.last_kernel_ms()
    fwd.append(f)
    tb.append(t)
e1.record()
torch.cuda.synchronize()
cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
ctx.ber_count(out, msgs, cnt)
ctx.sync()
fm, tm = sum(fwd) / n, sum(tb) / n
print(json.dumps({"frames": frames, "forward_ms": fm, "traceback_ms": tm, "gbit_s": frames * bits / ((fm + tm) * 1e-3) / 1e9,
                  "forward_only_gbit_s": frames * bits / (fm * 1e-3) / 1e9, "bit_errors": int(cnt[0]), "bits": int(cnt[1])}))

# the same channel output quantised to 3 bits: byte metrics, one byte per segment (ced_decode_batch_softq)
sigma_i8 = 32.0 * 10 ** (-3.0 / 20)
syms = ctx.quantize_soft(soft, T, 0.6 * sigma_i8, sym_stride=4112)
hard = ctx.decode_batch(ced.K7_DEFAULT, ctx.slice_soft_to_bytes(soft, T, seg_stride=4112), bits)
for _ in range(3):
    ctx.decode_batch_softq(ced.K7_DEFAULT, syms, bits, out=out)
fwd, tb = [], []
for _ in range(n):
    ctx.decode_batch_softq(ced.K7_DEFAULT, syms, bits, out=out)
    f, t = ctx.last_kernel_ms()
    fwd.append(f)
    tb.append(t)
torch.cuda.synchronize()
cnt = torch.zeros(4, dtype=torch.int64, device="cuda")
ctx.ber_count(out, msgs, cnt[:2])
ctx.ber_count(hard, msgs, cnt[2:])
ctx.sync()
fm, tm = sum(fwd) / n, sum(tb) / n
print(json.dumps({"softq": True, "frames": frames, "forward_ms": fm, "traceback_ms": tm,
                  "gbit_s": frames * bits / ((fm + tm) * 1e-3) / 1e9, "forward_only_gbit_s": frames * bits / (fm * 1e-3) / 1e9,
                  "bit_errors": int(cnt[0]), "hard_bit_errors": int(cnt[2]), "bits": int(cnt[1])}))
