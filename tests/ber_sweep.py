"""BASELINE config 4: BPSK+AWGN (hard-sliced => BSC) Eb/N0 sweep, GPU BER vs the reference's decoder.

For every Eb/N0 point the whole chain runs on the GPU through the C ABI
(random messages -> ced_encode_batch -> ced_bsc_channel -> ced_decode_batch -> ced_ber_count); the four
counters {channel flips, coded bits, decoded errors, decoded bits} stay on the device and are summed
over ranks with ONE all-reduce (NCCL when launched under torchrun).  "Identical to the reference"
is checked on a subset of each point: the very same noisy symbols are copied to the host, decoded by
the unmodified reference (oracle/_ref; the oracle port if it is absent) and the decoded-error counts
must be equal as integers.  berTestK7's own three points (berTestK7/berTestK7.c:95-96) lie on this
curve at Eb/N0 = 4.03 / 5.03 / 6.03 dB.

    python tests/ber_sweep.py [--frames-per-gpu N] [--subset M] [--out profiles/ber_sweep_r1.json]
    python -m torch.distributed.run --nproc-per-node 2 ... tests/ber_sweep.py

This file lives under tests/ because it uses the oracle as its checker.
"""
import argparse
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

PKT_BITS = 2048  # berTestK7.c:8 ENCODE_PKT_BYTE_LEN = 2048/8


def q_function(x):
    return 0.5 * math.erfc(x / math.sqrt(2.0))


def bsc_probability(ebn0_db, rate=0.5):
    """Hard-sliced BPSK over AWGN: p = Q(sqrt(2 * Rc * Eb/N0))."""
    return q_function(math.sqrt(2.0 * rate * 10.0 ** (ebn0_db / 10.0)))


def run_point_soft(ctx, code, ebn0_db, frames, seed, first_frame, subset, checker, soft_checker):
    """One Eb/N0 point on the AWGN channel with int8 soft output: the SAME channel realisation is decoded by the
    soft-decision decoder (ced_decode_batch_soft) and, after slicing to signs, by the hard-decision decoder.
    and, quantised to 3 bits (step 0.6 sigma), by the byte-metric soft decoder (ced_decode_batch_softq).
    counters = {sign errors, coded bits, hard decoded errors, bits, soft decoded errors, bits, 3-bit soft errors, bits}."""
    import numpy as np
    import torch
    T = PKT_BITS + code.S
    stride = (T + 15) // 16 * 16
    msgs = torch.empty((frames, PKT_BITS // 8), dtype=torch.uint8, device="cuda")
    segs = torch.zeros((frames, stride), dtype=torch.uint8, device="cuda")
    counters = torch.zeros(8, dtype=torch.int64, device="cuda")
    ctx.random_bytes(msgs, seed=seed, first_frame=first_frame)
    ctx.encode_batch(code, msgs, out=segs)
    soft = ctx.awgn_channel(segs, T, ebn0_db, seed=seed + 1, first_frame=first_frame, counters=counters[:2])
    dec_soft = ctx.decode_batch_soft(code, soft, PKT_BITS)
    hard = ctx.slice_soft_to_bytes(soft, T, seg_stride=stride)
    dec = ctx.decode_batch(code, hard, PKT_BITS)
    ctx.ber_count(dec, msgs, counters[2:4])
    ctx.ber_count(dec_soft, msgs, counters[4:6])
    sigma_i8 = 32.0 * 10.0 ** (-ebn0_db / 20.0)            # rate 1/2: sigma = 1 / sqrt(Eb/N0), amplitude 32
    syms = ctx.quantize_soft(soft, T, 0.6 * sigma_i8, sym_stride=stride)
    dec_q = ctx.decode_batch_softq(code, syms, PKT_BITS)
    ctx.ber_count(dec_q, msgs, counters[6:8])
    ctx.sync()
    check = None
    if subset and checker is not None:
        m, ms = min(subset, frames), min(max(subset // 10, 1), frames)
        want = checker(hard[:m, :T].cpu().numpy(), T)
        want_soft = soft_checker(soft[:ms, :2 * T].cpu().numpy(), T)
        check = {"frames": m, "bytes_identical": bool(np.array_equal(want, dec[:m].cpu().numpy())),
                 "reference_decoded_errors": int(np.bitwise_count(want ^ msgs[:m].cpu().numpy()).sum()),
                 "gpu_decoded_errors": int(np.bitwise_count((dec[:m] ^ msgs[:m]).cpu().numpy()).sum()),
                 "soft_frames": ms, "soft_bytes_identical": bool(np.array_equal(want_soft, dec_soft[:ms].cpu().numpy()))}
        sy = syms[:ms, :T].cpu().numpy()
        q8 = np.empty((ms, 2 * T), dtype=np.int8)
        q8[:, 0::2] = 7 - 2 * (sy & 7).astype(np.int16)
        q8[:, 1::2] = 7 - 2 * ((sy >> 3) & 7).astype(np.int16)
        check["softq_bytes_identical"] = bool(np.array_equal(soft_checker(q8, T), dec_q[:ms].cpu().numpy()))
    return counters, check


def crossing_db(rows, key, target):
    """Eb/N0 at which the BER curve rows[*][key] crosses `target` (log-linear interpolation), or None."""
    pts = [(r["ebn0_db"], r[key]) for r in rows if r[key] > 0]
    for (d0, b0), (d1, b1) in zip(pts, pts[1:]):
        if b0 >= target >= b1:
            return d0 + (d1 - d0) * (math.log(b0) - math.log(target)) / (math.log(b0) - math.log(b1))
    return None


def run_point(ctx, code, p, frames, seed, first_frame, subset, checker):
    import numpy as np
    import torch
    T = PKT_BITS + code.S
    stride = (T + 15) // 16 * 16
    msgs = torch.empty((frames, PKT_BITS // 8), dtype=torch.uint8, device="cuda")
    segs = torch.zeros((frames, stride), dtype=torch.uint8, device="cuda")
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    ctx.random_bytes(msgs, seed=seed, first_frame=first_frame)
    ctx.encode_batch(code, msgs, out=segs)
    ctx.bsc_channel(segs, T, code.n, p, seed=seed + 1, first_frame=first_frame, counters=counters[:2])
    dec = ctx.decode_batch(code, segs, PKT_BITS)
    ctx.ber_count(dec, msgs, counters[2:])
    ctx.sync()
    check = None
    if subset and checker is not None:
        m = min(subset, frames)
        noisy = segs[:m, :T].cpu().numpy()
        want = checker(noisy, T)
        ref_errs = int(np.bitwise_count(want ^ msgs[:m].cpu().numpy()).sum())
        gpu_errs = int(np.bitwise_count((dec[:m] ^ msgs[:m]).cpu().numpy()).sum())
        check = {"frames": m, "reference_decoded_errors": ref_errs, "gpu_decoded_errors": gpu_errs,
                 "bytes_identical": bool(np.array_equal(want, dec[:m].cpu().numpy()))}
    return counters, check


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames-per-gpu", type=int, default=1 << 18)
    ap.add_argument("--subset", type=int, default=10000, help="frames per point re-decoded by the reference (rank 0)")
    ap.add_argument("--points", default="0,1,2,3,4,5,6,7,8")
    ap.add_argument("--out", default="")
    ap.add_argument("--soft", action="store_true",
                    help="AWGN channel with int8 soft output: soft- and hard-decision decoding of the same noise")
    args = ap.parse_args(argv)

    import torch
    import torch.distributed as dist
    import convolutionalencdec_b200 as ced
    from convolutionalencdec_b200.sharding import allreduce_counts

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    checker, kind, soft_checker = None, None, None
    if rank == 0 and args.subset:
        import oracle
        soft_checker = lambda soft, T: oracle.port().decode_soft_batch(7, oracle.K7_G, soft, T)
        R = oracle.ref()
        if R is not None:
            checker, kind = (lambda noisy, T: R.decode_batch(noisy, T)), "reference (oracle/_ref)"
        else:
            P = oracle.port()
            checker, kind = (lambda noisy, T: P.decode_batch(7, oracle.K7_G, noisy, T)), "port (oracle/ced_oracle.c)"
    ctx = ced.Context(local_rank)
    code = ced.K7_DEFAULT
    from convolutionalencdec_b200 import ber_theory
    spectrum = ber_theory.distance_spectrum(7, [0o113, 0o171], 20)
    rows = []
    for i, db in enumerate(float(x) for x in args.points.split(",")):
        p = bsc_probability(db)
        if args.soft:
            counters, check = run_point_soft(ctx, code, db, args.frames_per_gpu, seed=1000 + 10 * i,
                                             first_frame=rank * args.frames_per_gpu,
                                             subset=args.subset if rank == 0 else 0, checker=checker,
                                             soft_checker=soft_checker)
            allreduce_counts(counters)
            c = [int(x) for x in counters.cpu().tolist()]
            rows.append({"ebn0_db": db, "bsc_p": p, "sign_errors": c[0], "coded_bits": c[1], "channel_ber": c[0] / c[1],
                         "hard_decoded_errors": c[2], "soft_decoded_errors": c[4], "decoded_bits": c[3],
                         "hard_ber": c[2] / c[3], "soft_ber": c[4] / c[5], "softq_decoded_errors": c[6],
                         "softq_ber": c[6] / c[7],
                         "uncoded_bpsk_ber": ber_theory.bpsk_ber(db), "subset_check": check})
            if rank == 0:
                print("Eb/N0 %4.1f dB  channel BER %.5e (Q: %.5e)  hard BER %.4e  soft BER %.4e  3-bit soft BER %.4e%s"
                      % (db, c[0] / c[1], p, c[2] / c[3], c[4] / c[5], c[6] / c[7],
                         "" if not check else "  subset: hard %d frames identical to the reference: %s, soft %d frames "
                         "identical to the soft oracle: %s, 3-bit soft: %s"
                         % (check["frames"], check["bytes_identical"], check["soft_frames"], check["soft_bytes_identical"],
                            check["softq_bytes_identical"])),
                      file=sys.stderr)
            continue
        counters, check = run_point(ctx, code, p, args.frames_per_gpu, seed=1000 + 10 * i,
                                    first_frame=rank * args.frames_per_gpu, subset=args.subset if rank == 0 else 0,
                                    checker=checker)
        allreduce_counts(counters)   # the only collective of BER mode: 4 x int64
        c = [int(x) for x in counters.cpu().tolist()]
        rows.append({"ebn0_db": db, "bsc_p": p, "channel_flips": c[0], "coded_bits": c[1], "decoded_errors": c[2],
                     "decoded_bits": c[3], "channel_ber": c[0] / c[1], "decoded_ber": c[2] / c[3],
                     "union_bound_hard": ber_theory.hard_decision_ber(p, spectrum),
                     "uncoded_bpsk_ber": ber_theory.bpsk_ber(db), "subset_check": check})
        if rank == 0:
            print("Eb/N0 %4.1f dB  p=%.5f  channel BER %.5e  decoded BER %.5e  (%d errors / %d bits)%s"
                  % (db, p, c[0] / c[1], c[2] / c[3], c[2], c[3],
                     "" if not check else "  subset %d frames: ref %d == gpu %d : %s"
                     % (check["frames"], check["reference_decoded_errors"], check["gpu_decoded_errors"],
                        check["bytes_identical"])), file=sys.stderr)
    if args.soft:
        hard_x, soft_x = crossing_db(rows, "hard_ber", 1e-4), crossing_db(rows, "soft_ber", 1e-4)
        softq_x = crossing_db(rows, "softq_ber", 1e-4)
        result = {"config": "K=7 r=1/2 g=(0113,0171), %d-bit packets, BPSK+AWGN quantised to int8 (amplitude 32): "
                            "soft-decision vs hard-decision decoding of the same channel output" % PKT_BITS,
                  "n_gpus": world, "frames_per_gpu": args.frames_per_gpu, "checker": kind,
                  "ebn0_db_at_ber_1e-4": {"hard": hard_x, "soft": soft_x, "soft_3bit": softq_x,
                                          "soft_decision_gain_db": (hard_x - soft_x) if hard_x and soft_x else None,
                                          "soft_3bit_gain_db": (hard_x - softq_x) if hard_x and softq_x else None},
                  "points": rows}
        if rank == 0:
            text = json.dumps(result, indent=1)
            if args.out:
                with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
                    f.write(text + "\n")
            print(json.dumps(result))
        ctx.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return result
    result = {"config": "K=7 r=1/2 g=(0113,0171), %d-bit packets, hard-decision BSC from BPSK+AWGN" % PKT_BITS,
              "distance_spectrum": {"dfree": spectrum.dfree, "weight": spectrum.weight, "event": spectrum.event},
              "n_gpus": world, "frames_per_gpu": args.frames_per_gpu, "checker": kind, "points": rows}
    if rank == 0:
        text = json.dumps(result, indent=1)
        if args.out:
            with open(os.path.join(ROOT, args.out) if not os.path.isabs(args.out) else args.out, "w") as f:
                f.write(text + "\n")
        print(json.dumps(result))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return result


if __name__ == "__main__":
    main()
