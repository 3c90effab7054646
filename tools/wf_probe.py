"""One call of each small-batch decode path per shape (for ncu: kernel names wfDecodeKernel / wsBlockKernel / wsJoinKernel)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

ctx = ced.Context(0)
rng = np.random.default_rng(1)
for frames, bits in ((1, 2048), (16, 2048), (1, 4096), (148, 4096)):
    msgs = torch.from_numpy(rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)).cuda()
    segs = torch.zeros((frames, bits + 16), dtype=torch.uint8, device="cuda")
    ctx.encode_batch(ced.K7_DEFAULT, msgs, out=segs)
    ctx.bsc_channel(segs, bits + 6, 2, 0.0377, seed=2)
    for split in ("1", "0"):
        os.environ["CED_WARP_SPLIT"] = split
        for _ in range(2):
            out = ctx.decode_batch(ced.K7_DEFAULT, segs, bits)
        ctx.sync()
ctx.close()
