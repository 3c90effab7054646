"""e2e rate of ced_decode_batch_host on page-locked buffers for the transfer-compression modes
(CED_HOST_PACK = 0 raw copies, 1 every chunk packed by host threads, 2 adaptive: only chunks the copy engine is
not ready for), from 1 and 2 host threads.  Usage: host_pack_modes.py [threads_per_context ...]"""
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

frames, bits = 1 << 16, 4096
code = ced.K7_DEFAULT
pool_sizes = [int(a) for a in sys.argv[1:]] or [8]
for pool in pool_sizes:
    os.environ["CED_HOST_THREADS"] = str(pool)
    ctxs = [ced.Context(0), ced.Context(0)]
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctxs[0].random_bytes(msgs, seed=1)
    segs = ctxs[0].encode_batch(code, msgs, seg_stride=4112)
    ctxs[0].bsc_channel(segs, bits + 6, 2, 0.0377, seed=1)
    want = ctxs[0].decode_batch(code, segs, bits)
    ctxs[0].sync()
    h_in = [segs.cpu().pin_memory() for _ in ctxs]
    h_out = [torch.empty((frames, bits // 8), dtype=torch.uint8).pin_memory() for _ in ctxs]

    def rate(n_threads, n=6):
        def worker(i):
            for _ in range(n):
                ctxs[i].decode_batch_host(code, h_in[i], bits, h_out[i])
        for i in range(n_threads):
            ctxs[i].decode_batch_host(code, h_in[i], bits, h_out[i])
        t0 = time.perf_counter()
        ts = [threading.Thread(target=worker, args=(i,)) for i in range(n_threads)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        el = time.perf_counter() - t0
        for i in range(n_threads):
            assert torch.equal(h_out[i].cuda(), want)
        return frames * bits * n * n_threads / el / 1e9

    modes = [m.split(":") for m in os.environ.get("MODES", "0:1,1:1,2:1,2:2,2:3").split(",")]
    for mode, look in modes:
        os.environ["CED_HOST_PACK"] = mode
        os.environ["CED_HOST_PACK_LOOKBACK"] = look
        print("pool %2d  CED_HOST_PACK=%s lookback=%s : 1 caller %.1f Gbit/s, 2 callers %.1f Gbit/s"
              % (pool, mode, look, rate(1), rate(2)), flush=True)
    for c in ctxs:
        c.close()
