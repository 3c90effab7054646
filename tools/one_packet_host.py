import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import convolutionalencdec_b200 as ced
ctx = ced.Context(0)
rng = np.random.default_rng(1)
for bits in (2048, 4096, 16384):
    T = bits + 6
    msgs = rng.integers(0, 256, (1, bits // 8), dtype=np.uint8)
    segs = np.zeros((1, T), dtype=np.uint8)
    ctx.encode_batch_host(ced.K7_DEFAULT, msgs, segs)
    out = np.zeros((1, bits // 8), dtype=np.uint8)
    for pinned in (False, True):
        s2, o2 = segs, out
        if pinned:
            s2 = torch.from_numpy(segs.copy()).pin_memory().numpy(); o2 = torch.from_numpy(out.copy()).pin_memory().numpy()
        for _ in range(20):
            ctx.decode_batch_host(ced.K7_DEFAULT, s2, bits, o2)
        assert np.array_equal(o2, msgs)
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 0.5:
            ctx.decode_batch_host(ced.K7_DEFAULT, s2, bits, o2); n += 1
        print(bits, "pinned" if pinned else "pageable", "%.1f us per call" % ((time.perf_counter() - t0) / n * 1e6))
ctx.close()
