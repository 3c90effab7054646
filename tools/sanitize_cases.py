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
# per-frame API
api = ced.RefApi("k7")
enc = api.encoder(); enc.resetConvEncoder(); enc.initConvEncoder()
dec = api.decoder(); dec.VITERBI_RESET(); dec.VITERBI_INIT()
msg = np.arange(64, dtype=np.uint8)
s = enc.convEnc(msg, True)
dec.VITERBI_DECODER_HARD(s[:100], False)
out = dec.VITERBI_DECODER_HARD(s[100:], True)
ok &= bool(np.array_equal(out, msg))
print("per-frame chunked round trip:", bool(np.array_equal(out, msg)))
ctx.close()
print("ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
