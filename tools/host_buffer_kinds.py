"""e2e rate of ced_decode_batch_host for the kinds of host buffers a caller may own: page-locked (pinned),
ordinary pageable memory, and pageable memory registered on the fly with cudaHostRegister."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

frames, bits = 1 << 16, 4096
T = bits + 6
ctx = ced.Context(0)
code = ced.K7_DEFAULT
msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
ctx.random_bytes(msgs, seed=1)
segs = ctx.encode_batch(code, msgs, seg_stride=4112)
ctx.sync()
pageable_in = segs.cpu().numpy().copy()
pageable_out = np.zeros((frames, bits // 8), dtype=np.uint8)
pinned_in = torch.from_numpy(pageable_in).pin_memory()
pinned_out = torch.empty((frames, bits // 8), dtype=torch.uint8).pin_memory()


def rate(i, o, n=5):
    ctx.decode_batch_host(code, i, bits, o)
    t0 = time.perf_counter()
    for _ in range(n):
        ctx.decode_batch_host(code, i, bits, o)
    return frames * bits * n / (time.perf_counter() - t0) / 1e9


print("pinned   : %.1f Gbit/s" % rate(pinned_in, pinned_out))
print("pageable : %.1f Gbit/s (worker threads pack to 2 bits into page-locked staging)" % rate(pageable_in, pageable_out))
assert np.array_equal(pageable_out, msgs.cpu().numpy())
pageable_out[:] = 0
os.environ["CED_HOST_PACK"] = "0"
print("pageable, direct copies (CED_HOST_PACK=0): %.1f Gbit/s" % rate(pageable_in, pageable_out))
assert np.array_equal(pageable_out, msgs.cpu().numpy())
del os.environ["CED_HOST_PACK"]
t0 = time.perf_counter()
ctx.host_register(pageable_in)
ctx.host_register(pageable_out)
reg = time.perf_counter() - t0
print("registered on the fly: %.1f Gbit/s (+ %.1f ms once to register %.0f MB)"
      % (rate(pageable_in, pageable_out), reg * 1e3, (pageable_in.nbytes + pageable_out.nbytes) / 1e6))
ctx.host_unregister(pageable_in)
ctx.host_unregister(pageable_out)
ctx.close()
