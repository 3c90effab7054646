"""Randomised parity soak: random codes, frame shapes, strides, base offsets and channels through every batched decode entry
point, each result compared with the CPU oracle (test infrastructure, like tests/).  Stops at the first difference.
    python tools/fuzz_parity.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import convolutionalencdec_b200 as ced  # noqa: E402
import oracle  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(seed)
P = oracle.port()
ctx = ced.Context(0)
K7 = oracle.K7_G


def rand_gen(K, rng, both_ends):
    g = int(rng.integers(1, 1 << K))
    if both_ends:
        g |= 1 | (1 << (K - 1))
    return g


def noisy(clean, n, p):
    flips = rng.random(clean.shape + (n,)) < p
    out = clean.copy()
    for j in range(n):
        out ^= (flips[..., j].astype(np.uint8) << j)
    return out


def place(arr, pad, off, fill=0xEE):
    frames, T = arr.shape
    flat = torch.full((frames * (T + pad) + 64,), fill, dtype=torch.uint8, device="cuda")
    view = flat[off:off + frames * (T + pad)].view(frames, T + pad)
    view[:, :T] = torch.from_numpy(arr).cuda()
    return view


counts = {}
t_end = time.time() + seconds
it = 0
while time.time() < t_end:
    it += 1
    kind = rng.choice(["k7", "k7rt", "generic", "packed", "soft", "softq", "k2", "window", "windowq", "windowgen", "enc", "host"])
    frames = int(rng.choice([1, 2, 31, 33, 64, 100, 257]))
    bits = int(rng.choice([8, 16, 40, 96, 104, 200, 512, 1000, 2048]))
    pad, off = int(rng.integers(0, 20)), int(rng.integers(0, 16))
    p = float(rng.choice([0.0, 0.02, 0.06, 0.2, 0.5]))
    msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
    tag = kind
    # small batches of K = 6, 7 codes: the warp-per-frame kernel (the library's default) or the thread-per-frame kernels
    if rng.integers(0, 2):
        os.environ["CED_WARP_FRAME_MAX"] = "0"
    else:
        os.environ.pop("CED_WARP_FRAME_MAX", None)
    os.environ["CED_WARP_FRAME_RADIX"] = str(int(rng.choice([2, 4])))
    os.environ["CED_WARP_SPLIT"] = str(int(rng.integers(0, 2)))             # frames cut in time as well (warp_split.cu)
    os.environ["CED_WARP_SPLIT_WARMUP"] = str(int(rng.choice([8, 48, 96])))
    if kind == "enc":
        k = int(rng.choice([1, 1, 2]))
        K = int(rng.integers(2, 10)) if k == 1 else int(rng.integers(2, 6))
        n = int(rng.integers(1, 9)) if k == 1 else int(rng.integers(2, 5))
        g = [int(rng.integers(1, 1 << (k * K))) for _ in range(n)]
        code = ced.Code(K, g)
        d_msgs = torch.from_numpy(msgs).cuda()
        if k == 1:
            want = P.encode_batch(K, g, msgs)
            T = want.shape[1]
            out_t = torch.full((frames, T + pad), 0xEE, dtype=torch.uint8, device="cuda")
            ctx.encode_batch(code, d_msgs, out=out_t)
        else:
            want = P.encode_batch_k(K, 2, g, msgs)
            T = want.shape[1]
            out_t = torch.full((frames, T + pad), 0xEE, dtype=torch.uint8, device="cuda")
            ctx.encode_batch_k(code, 2, d_msgs, out=out_t)
        ctx.sync()
        got = out_t[:, :T]
        if pad and not bool((out_t[:, T:] == 0xEE).all()):
            print("ENCODER WROTE PAST THE ROW", dict(K=K, k=k, n=n, frames=frames, bits=bits, pad=pad))
            sys.exit(1)
        tag = "enc k=%d" % k
    elif kind == "host":
        g = [K7, [0o133, 0o171]][int(rng.integers(0, 2))]
        code = ced.Code(7, g)
        T = bits + 6
        frames = int(rng.choice([1, 100, 9000]))
        msgs = rng.integers(0, 256, (frames, bits // 8), dtype=np.uint8)
        rx = noisy(P.encode_batch(7, g, msgs), 2, min(p, 0.06))
        h_in = np.full((frames, T + pad), 0xEE, dtype=np.uint8)
        h_in[:, :T] = rx
        h_out = np.zeros((frames, bits // 8), dtype=np.uint8)
        ctx.decode_batch_host(code, h_in, bits, h_out)
        want = P.decode_batch(7, g, rx[:200], T)
        got = torch.from_numpy(h_out[:200])
        enc_out = np.zeros((frames, T + pad), dtype=np.uint8)
        ctx.encode_batch_host(code, msgs, enc_out)
        if not np.array_equal(enc_out[:200, :T], P.encode_batch(7, g, msgs[:200])):
            print("HOST ENCODE MISMATCH", dict(frames=frames, bits=bits, pad=pad))
            sys.exit(1)
    elif kind == "windowq":
        g = [K7, [0o133, 0o171]][int(rng.integers(0, 2))]
        code = ced.Code(7, g)
        total_bits = int(rng.choice([320, 1000, 2048]))
        Tt = total_bits + 6
        frames = min(frames, 64)
        x = rng.integers(0, 8, (frames, Tt, 2))
        syms = (x[..., 0] | (x[..., 1] << 3)).astype(np.uint8)
        s8 = np.empty((frames, 2 * Tt), dtype=np.int8)
        s8[:, 0::2] = 7 - 2 * (syms & 7).astype(np.int16)
        s8[:, 1::2] = 7 - 2 * ((syms >> 3) & 7).astype(np.int16)
        call, depth = int(rng.choice([96, 192, 480])), int(rng.choice([24, 48, 96]))
        wd = ctx.window_decoder(code, frames, depth=depth, softq=True)
        d = torch.from_numpy(syms).cuda()
        parts = [wd.push(d[:, a:min(a + call, Tt)], last=a + call >= Tt).clone() for a in range(0, Tt, call)]
        got = torch.cat(parts, dim=1)
        want = np.stack([P.decode_window_soft(7, g, s8[i], call, depth) for i in range(frames)])
    elif kind in ("k7", "packed", "soft", "softq", "window"):
        g = [K7, [0o133, 0o171], [0o171, 0o133]][int(rng.integers(0, 3 if kind in ("k7", "packed") else 2))]
        code = ced.Code(7, g)
        T = bits + 6
        clean = P.encode_batch(7, g, msgs)
        rx = noisy(clean, 2, p)
        if kind == "k7":
            want = P.decode_batch(7, g, rx, T)
            got = ctx.decode_batch(code, place(rx | (rng.integers(0, 64, rx.shape, dtype=np.uint8) << 2), pad, off), bits)
        elif kind == "packed":
            want = P.decode_batch(7, g, rx, T)
            pk = ctx.pack_symbols(torch.from_numpy(rx).cuda(), T)
            got = ctx.decode_batch_packed(code, pk, bits)
        elif kind == "soft":
            s = rng.integers(-128, 128, (frames, 2 * T), dtype=np.int8)
            sign = 1 - 2 * ((rx[..., None] >> np.arange(2)) & 1).astype(np.int16)
            mag = rng.integers(0, 100, rx.shape + (2,))
            s = np.clip(sign * mag, -128, 127).astype(np.int8).reshape(frames, 2 * T)
            stride = (2 * T + 15) // 16 * 16
            d = torch.zeros((frames, stride), dtype=torch.int8, device="cuda")
            d[:, :2 * T] = torch.from_numpy(s).cuda()
            want = P.decode_soft_batch(7, g, s, T)
            got = ctx.decode_batch_soft(code, d, bits)
        elif kind == "softq":
            x = rng.integers(0, 8, rx.shape + (2,))
            keep = rng.random(rx.shape + (2,)) > p          # mostly the right side of 3.5
            bit = (rx[..., None] >> np.arange(2)) & 1
            x = np.where(keep, np.where(bit == 1, 4 + x // 2, 3 - x // 2), x)
            syms = (x[..., 0] | (x[..., 1] << 3)).astype(np.uint8)
            s = np.empty((frames, 2 * T), dtype=np.int8)
            s[:, 0::2] = 7 - 2 * (syms & 7).astype(np.int16)
            s[:, 1::2] = 7 - 2 * ((syms >> 3) & 7).astype(np.int16)
            want = P.decode_soft_batch(7, g, s, T)
            got = ctx.decode_batch_softq(code, place(syms | (rng.integers(0, 4, syms.shape, dtype=np.uint8) << 6), pad, off), bits)
        else:
            total_bits = int(rng.choice([320, 1000, 2048]))
            msgs = rng.integers(0, 256, (frames, total_bits // 8), dtype=np.uint8)
            clean = P.encode_batch(7, g, msgs)
            rx = noisy(clean, 2, min(p, 0.06))
            Tt = total_bits + 6
            call, depth = int(rng.choice([96, 192, 480])), int(rng.choice([24, 48, 96]))
            wd = ctx.window_decoder(code, frames, depth=depth)
            d = torch.from_numpy(rx).cuda()
            parts = [wd.push(d[:, a:min(a + call, Tt)], last=a + call >= Tt).clone() for a in range(0, Tt, call)]
            got = torch.cat(parts, dim=1)
            want = np.stack([P.decode_window(7, g, rx[i], call, depth) for i in range(frames)])
            bits = total_bits
    elif kind == "windowgen":
        K = int(rng.choice([3, 4, 5, 7]))
        n = int(rng.integers(2, 4))
        g = [rand_gen(K, rng, False) for _ in range(n)]
        code = ced.Code(K, g)
        total_bits = int(rng.choice([320, 1000, 2048]))
        msgs = rng.integers(0, 256, (frames, total_bits // 8), dtype=np.uint8)
        rx = noisy(P.encode_batch(K, g, msgs), n, min(p, 0.2))
        Tt = total_bits + K - 1
        call, depth = int(rng.choice([96, 192, 480])), int(rng.choice([24, 48, 96]))
        wd = ctx.window_decoder(code, frames, depth=depth)
        d = place(rx, pad, off)
        parts = [wd.push(d[:, a:min(a + call, Tt)], last=a + call >= Tt).clone() for a in range(0, Tt, call)]
        got = torch.cat(parts, dim=1)
        want = np.stack([P.decode_window(K, g, rx[i], call, depth, symmetric=False) for i in range(frames)])
        bits = total_bits
        tag = "windowgen K=%d" % K
    elif kind == "k7rt":
        n = int(rng.integers(2, 4))
        g = [rand_gen(7, rng, True) for _ in range(n)]
        code = ced.Code(7, g)
        T = bits + 6
        rx = noisy(P.encode_batch(7, g, msgs), n, p)
        want = P.decode_batch(7, g, rx, T, symmetric=False)
        got = ctx.decode_batch(code, place(rx, pad, off), bits)
    elif kind == "generic":
        K = int(rng.integers(3, 10))
        n = int(rng.integers(2, 4))
        g = [rand_gen(K, rng, False) for _ in range(n)]
        code = ced.Code(K, g)
        T = bits + K - 1
        rx = noisy(P.encode_batch(K, g, msgs), n, p)
        want = P.decode_batch(K, g, rx, T, symmetric=False)
        got = ctx.decode_batch(code, place(rx, pad, off), bits)
        tag = "generic K=%d" % K
    else:
        K = int(rng.integers(2, 6))
        n = int(rng.integers(2, 5))
        g = [int(rng.integers(1, 1 << (2 * K))) for _ in range(n)]
        code = ced.Code(K, g)
        clean = P.encode_batch_k(K, 2, g, msgs)
        T = clean.shape[1]
        rx = noisy(clean, n, p)
        want = P.decode_batch_k(K, 2, g, rx, T)
        got = ctx.decode_batch_k(code, 2, place(rx, pad, off), bits)
        tag = "k2 K=%d n=%d" % (K, n)
    ctx.sync()
    if not np.array_equal(got.cpu().numpy()[:, :want.shape[1]], want):
        print("MISMATCH", kind, dict(frames=frames, bits=bits, pad=pad, off=off, p=p, g=[oct(x) for x in g], seed=seed, it=it))
        sys.exit(1)
    counts[tag] = counts.get(tag, 0) + 1
print("fuzz ok: %d cases in %.0f s: %s" % (it, seconds, dict(sorted(counts.items()))))
ctx.close()
