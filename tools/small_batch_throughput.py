"""Small batches: ced_decode_batch through the warp-per-frame kernel (csrc/warp_frame.cu) and through the
thread-per-frame kernels (CED_WARP_FRAME_MAX=0), device-resident, CUDA events; and speedDecode's shape -- 16 packets of
2048 bits -- as one ced_decode_batch_host call against 16 synchronous per-packet calls (speedDecode.c:78-79).
usage: python tools/small_batch_throughput.py [reps]"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import convolutionalencdec_b200 as ced  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
INNER = 8
ctx = ced.Context(0)
code = ced.K7_DEFAULT
rng = np.random.default_rng(1)


def timed(fn, reps):
    for _ in range(3):
        fn()
    ctx.sync()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        a.record()
        for _ in range(INNER):   # back to back: the launch latency of all but the first call hides behind the previous one
            fn()
        b.record()
        b.synchronize()
        ts.append(a.elapsed_time(b) * 1e3 / INNER)
    return float(np.median(ts))


print("frames  bits   thread-per-frame us   warp-per-frame us (radix 2)   (radix 4)   cut in time   gain   Gbit/s")
for bits in (2048, 4096):
    for frames in (1, 16, 64, 148, 296, 592, 1024, 2048, 4096):
        msgs = torch.from_numpy(rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)).cuda()
        segs = torch.zeros((frames, (bits + 6 + 15) // 16 * 16), dtype=torch.uint8, device="cuda")
        ctx.encode_batch(code, msgs, out=segs)
        ctx.bsc_channel(segs, bits + 6, 2, 0.0377, seed=2)
        out = torch.zeros((frames, bits // 8), dtype=torch.uint8, device="cuda")
        res = {}
        for name, env in (("tpf", {"CED_WARP_FRAME_MAX": "0", "CED_WARP_SPLIT": "0"}),
                          ("r2", {"CED_WARP_FRAME_MAX": "100000", "CED_WARP_FRAME_RADIX": "2"}),
                          ("r4", {"CED_WARP_FRAME_MAX": "100000", "CED_WARP_FRAME_RADIX": "4"}),
                          ("split", {"CED_WARP_SPLIT": "1"})):
            os.environ.update(env)
            res[name] = timed(lambda: ctx.decode_batch(code, segs, bits, out=out), reps)
            res[name + "_out"] = out.clone()
        assert torch.equal(res["tpf_out"], res["r2_out"]) and torch.equal(res["tpf_out"], res["r4_out"])
        assert torch.equal(res["tpf_out"], res["split_out"])
        best = min(res["r4"], res["split"])
        print("%6d %5d %12.1f %22.1f %18.1f %12.1f %9.2f %8.2f" % (frames, bits, res["tpf"], res["r2"], res["r4"], res["split"],
                                                                    res["tpf"] / best, frames * bits / best / 1e3))
for k in ("CED_WARP_FRAME_MAX", "CED_WARP_FRAME_RADIX", "CED_WARP_SPLIT"):
    os.environ.pop(k, None)

# speedDecode's shape from host buffers
bits, T = 2048, 2054
api = ced.RefApi("k7")
enc = api.encoder(); enc.resetConvEncoder(); enc.initConvEncoder()
dec = api.decoder(); dec.VITERBI_RESET(); dec.VITERBI_INIT()
msgs = rng.integers(0, 256, (16, bits // 8), dtype=np.uint8)
segs = np.stack([enc.convEnc(m, True) for m in msgs])[:, :T].copy()
u8p = ctypes.POINTER(ctypes.c_uint8)
out1 = np.zeros(bits // 8 + 8, dtype=np.uint8)
ptrs, outp = [segs[i].ctypes.data_as(u8p) for i in range(16)], out1.ctypes.data_as(u8p)
call = api.lib.viterbiDecoderHardButterflyk1
for i in range(32):
    call(dec.p, ptrs[i % 16], outp, T, True)
n, t0 = 0, time.perf_counter()
while time.perf_counter() - t0 < 1.0:
    for i in range(16):
        call(dec.p, ptrs[i], outp, T, True)
    n += 1
per_packet_us = (time.perf_counter() - t0) / n * 1e6
outb = np.zeros((16, bits // 8), dtype=np.uint8)
for label, env in (("thread-per-frame kernels", "0"), ("warp-per-frame kernel", None)):
    if env is None:
        os.environ.pop("CED_WARP_FRAME_MAX", None)
    else:
        os.environ["CED_WARP_FRAME_MAX"] = env
    for _ in range(10):
        ctx.decode_batch_host(code, segs, bits, outb)
    assert np.array_equal(outb, msgs)
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < 1.0:
        ctx.decode_batch_host(code, segs, bits, outb)
        n += 1
    batch_us = (time.perf_counter() - t0) / n * 1e6
    print("16 x 2048-bit packets from host buffers: 16 synchronous per-packet calls %.1f us; one ced_decode_batch_host call "
          "(%s) %.1f us = %.2f x" % (per_packet_us, label, batch_us, per_packet_us / batch_us))
ctx.close()
