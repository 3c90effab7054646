import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import convolutionalencdec_b200 as ced
ctx = ced.Context(0)
rng = np.random.default_rng(1)
frames, bits = 148, 4096
msgs = torch.from_numpy(rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)).cuda()
segs = torch.zeros((frames, 4112), dtype=torch.uint8, device="cuda")
ctx.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
ctx.bsc_channel(segs, bits + 6, 2, 0.0377, seed=2)
for r in ("4", "2"):
    os.environ["CED_WARP_FRAME_RADIX"] = r
    out = ctx.decode_batch(ced.K7_DEFAULT, segs, bits)
    ctx.sync()
ctx.close()
