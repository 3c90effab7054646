"""Decode `frames` x 4096-bit frames a few times through ced_decode_batch (fused kernel when enabled) and print the
CUDA-event time per decode; run under ncu --metrics ... to see DRAM traffic vs. the size of the decision rings.
python tools/fused_probe.py frames [p]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import convolutionalencdec_b200 as ced

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 16
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0377
bits, T = 4096, 4102
ctx = ced.Context(0)
msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
ctx.random_bytes(msgs, seed=314)
segs = torch.zeros((frames, 4112), dtype=torch.uint8, device="cuda")
ctx.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
ctx.bsc_channel(segs, T, 2, p, seed=2718)
out = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
for _ in range(3):
    ctx.decode_batch(ced.K7_DEFAULT, segs, bits, out=out)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    ctx.decode_batch(ced.K7_DEFAULT, segs, bits, out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
back = ctx.last_fallback_frames()
ok = bool(torch.equal(out, msgs)) if p == 0 else None
cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
ctx.ber_count(out, msgs, cnt)
ctx.sync()
print(json.dumps({"frames": frames, "p": p, "ms": ms, "gbit_s": frames * bits / (ms * 1e-3) / 1e9, "handed_back": back,
                  "bit_errors": int(cnt[0]), "fused": os.environ.get("CED_FUSED", "1")}))
