"""A handful of small, awkwardly shaped calls through every kernel; meant to be run under
`compute-sanitizer --tool memcheck` (and racecheck) on the GPU box.  Checks results against
nothing but self-consistency (encode -> decode round trip) so it needs no oracle."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402

ctx = ced.Context(0)
code = ced.K7_DEFAULT
ok = True
for bits, frames, pad, off in ((8, 1, 0, 0), (96, 33, 0, 0), (104, 65, 3, 1), (512, 100, 10, 0), (1000 // 8 * 8, 37, 0, 5),
                               (4096, 40, 0, 0), (4096, 33, 10, 0)):
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=bits)
    stride = T + pad
    flat = torch.zeros(frames * stride + off, dtype=torch.uint8, device="cuda")   # exactly sized: no slack
    segs = flat[off:off + frames * stride].view(frames, stride)
    ctx.encode_batch(code, msgs, out=segs)
    ctx.bsc_channel(segs, T, 2, 0.01, seed=1)
    dec = ctx.decode_batch(code, segs, bits)
    packed = ctx.pack_symbols(segs, T)
    dec_p = ctx.decode_batch_packed(code, packed, bits)
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    ctx.ber_count(dec, msgs, cnt)
    ctx.sync()
    same = bool(torch.equal(dec, dec_p))
    errs = int(cnt[0])
    print("bits %5d frames %4d pad %2d off %d: packed==byte %s, bit errors %d" % (bits, frames, pad, off, same, errs))
    ok &= same and errs < bits * frames // 50 + 8
    pk2 = ctx.encode_batch_packed(code, msgs)
    ctx.sync()
# generic codes
for K, g in ((3, (7, 6)), (9, (0o561, 0o753))):
    c2 = ced.Code(K, g)
    msgs = torch.empty((50, 32), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=K)
    segs = ctx.encode_batch(c2, msgs)
    dec = ctx.decode_batch(c2, segs, 256)
    ctx.sync()
    ok &= bool(torch.equal(dec, msgs))
    print("K=%d generic round trip:" % K, bool(torch.equal(dec, msgs)))
# run-time generators on the SWAR kernel (n = 2 and 3) and the windowed decoder, exactly sized buffers
for g in ((0o171, 0o133), (0o133, 0o171, 0o165)):
    c3 = ced.Code(7, g)
    frames, bits = 45, 96 * 5 + 88
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=sum(g))
    flat = torch.zeros(frames * T + 3, dtype=torch.uint8, device="cuda")
    segs = flat[3:].view(frames, T)
    ctx.encode_batch(c3, msgs, out=segs)
    dec = ctx.decode_batch(c3, segs, bits)
    wd = ctx.window_decoder(c3, frames, depth=24)
    parts = [wd.push(segs[:, a:min(a + 192, T)], last=a + 192 >= T).clone() for a in range(0, T, 192)]
    ctx.sync()
    good = bool(torch.equal(dec, msgs)) and bool(torch.equal(torch.cat(parts, dim=1), msgs))
    ok &= good
    print("run-time code %s batch + windowed round trip:" % (g,), good)
# soft symbols
soft = torch.randint(-128, 128, (20, 2 * 262 + 4), dtype=torch.int8, device="cuda")
pk = ctx.slice_soft_symbols(soft, 262)
ctx.decode_batch_packed(code, pk, 256)
ctx.sync()
# round 2: true soft-decision decoder, AWGN channel, table-driven SIMD-in-word kernels for other codes, k > 1 codes,
# packed windowed decoder, fused kernels, resident packet decoder -- exactly sized buffers again
for frames, bits in ((1, 8), (33, 96), (70, 1000 // 8 * 8)):
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=frames)
    segs = ctx.encode_batch(code, msgs, seg_stride=(T + 15) // 16 * 16)
    softb = ctx.awgn_channel(segs, T, 9.0, seed=7)       # 9 dB: practically noise-free, rows of exactly ceil16(2 T) bytes
    d = ctx.decode_batch_soft(code, softb, bits)
    ctx.sync()
    good = bool(torch.equal(d, msgs))
    ok &= good
    print("soft-decision round trip %d x %d:" % (frames, bits), good)
    flatq = torch.zeros(frames * T + 2, dtype=torch.uint8, device="cuda")
    symq = flatq[2:].view(frames, T)                     # misaligned, exactly sized
    ctx.quantize_soft(softb, T, 8.0, out=symq)
    dq = ctx.decode_batch_softq(code, symq, bits)
    ctx.sync()
    goodq = bool(torch.equal(dq, msgs))
    ok &= goodq
    print("3-bit soft round trip %d x %d:" % (frames, bits), goodq)
for K, g in ((3, (7, 5, 3)), (4, (0o15, 0o17)), (5, (0o23, 0o35)), (7, (0o133, 0o170)), (9, (0o557, 0o663, 0o711))):
    c4 = ced.Code(K, g)
    frames, bits = 67, 200
    T = bits + K - 1
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=K + 50)
    flat = torch.zeros(frames * T + 1, dtype=torch.uint8, device="cuda")
    segs = flat[1:].view(frames, T)                      # misaligned base, no slack
    ctx.encode_batch(c4, msgs, out=segs)
    d = ctx.decode_batch(c4, segs, bits)
    ctx.sync()
    good = bool(torch.equal(d, msgs))
    ok &= good
    print("table-driven SWAR K=%d n=%d round trip:" % (K, len(g)), good)
for K, k, g in ((3, 2, (0o27, 0o75, 0o72)), (5, 2, (0o1236, 0o0155, 0o1337, 0o1701)), (2, 4, (0o357, 0o261, 0o173, 0o225, 0o316))):
    c5 = ced.Code(K, g)
    msgs = torch.empty((41, 24), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=K * 10 + k)
    segs = ctx.encode_batch_k(c5, k, msgs)
    d = ctx.decode_batch_k(c5, k, segs, 192)
    ctx.sync()
    good = bool(torch.equal(d, msgs))
    ok &= good
    print("k=%d K=%d round trip:" % (k, K), good)
for mode in ("1", "2"):
    os.environ["CED_FUSED"], os.environ["CED_FUSED_MIN_FRAMES"] = mode, "1"
    frames, bits = 130, 96 * 6 + 40
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=11)
    segs = ctx.encode_batch(code, msgs, seg_stride=(T + 15) // 16 * 16)
    d = ctx.decode_batch(code, segs, bits)
    ctx.sync()
    good = bool(torch.equal(d, msgs))
    ok &= good
    print("fused kernel mode %s round trip:" % mode, good)
os.environ.pop("CED_FUSED")
os.environ.pop("CED_FUSED_MIN_FRAMES")
# per-frame API
os.environ["CED_STREAM_SERVER"] = "1"
apis = ced.RefApi("k7")
encs = apis.encoder(); encs.resetConvEncoder(); encs.initConvEncoder()
decs = apis.decoder(); decs.VITERBI_RESET(); decs.VITERBI_INIT()
for nbytes in (1, 17, 256):
    m = np.arange(nbytes, dtype=np.uint8)
    good = bool(np.array_equal(decs.VITERBI_DECODER_HARD(encs.convEnc(m, True), True, max_bytes=4096), m))
    ok &= good
    print("resident packet decoder, %d bytes:" % nbytes, good)
os.environ["CED_STREAM_SERVER"] = "0"
api = ced.RefApi("k7")
enc = api.encoder(); enc.resetConvEncoder(); enc.initConvEncoder()
dec = api.decoder(); dec.VITERBI_RESET(); dec.VITERBI_INIT()
msg = np.arange(64, dtype=np.uint8)
s = enc.convEnc(msg, True)
dec.VITERBI_DECODER_HARD(s[:100], False)
out = dec.VITERBI_DECODER_HARD(s[100:], True)
ok &= bool(np.array_equal(out, msg))
print("per-frame chunked round trip:", bool(np.array_equal(out, msg)))
# small batches: the warp-per-frame kernel (radix 4 and 2) and the time-split kernels, exactly sized and misaligned
# input rows, output rows between canary bytes; the same calls on the thread-per-frame kernels as the yardstick
for frames, bits, pad, off in ((1, 8, 0, 0), (3, 2048, 0, 5), (16, 2048, 3, 0), (33, 4096, 1, 9), (5, 16384, 0, 1)):
    T = bits + 6
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    ctx.random_bytes(msgs, seed=bits + frames)
    stride = T + pad
    flat = torch.zeros(frames * stride + off, dtype=torch.uint8, device="cuda")
    segs = flat[off:off + frames * stride].view(frames, stride)
    ctx.encode_batch(code, msgs, out=segs)
    ctx.bsc_channel(segs, T, 2, 0.03, seed=2)
    os.environ["CED_WARP_FRAME_MAX"] = "0"
    want = ctx.decode_batch(code, segs, bits).clone()
    ctx.sync()
    os.environ.pop("CED_WARP_FRAME_MAX")
    for mode in ({"CED_WARP_SPLIT": "0", "CED_WARP_FRAME_RADIX": "4"}, {"CED_WARP_SPLIT": "0", "CED_WARP_FRAME_RADIX": "2"},
                 {"CED_WARP_SPLIT": "1"}, {"CED_WARP_SPLIT": "1", "CED_WARP_SPLIT_WARMUP": "8", "CED_WARP_SPLIT_LEN": "16"}):
        os.environ.update(mode)
        ostride = bits // 8 + 5
        guard = torch.full((frames * ostride + 64,), 0xA5, dtype=torch.uint8, device="cuda")
        out = guard[32:32 + frames * ostride].view(frames, ostride)
        ctx.decode_batch(code, segs, bits, out=out)
        ctx.sync()
        good = bool(torch.equal(out[:, :bits // 8], want)) and bool((out[:, bits // 8:] == 0xA5).all()) and \
            bool((guard[:32] == 0xA5).all()) and bool((guard[32 + frames * ostride:] == 0xA5).all())
        ok &= good
        print("small batch %3d x %5d bits, %s:" % (frames, bits, mode), good)
        for k in mode:
            os.environ.pop(k)
ctx.close()
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
