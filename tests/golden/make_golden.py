"""Generates tests/golden/k7_reference_vectors.npz from the UNMODIFIED reference
(oracle/_ref, i.e. /root/reference/src compiled by oracle/Makefile).  Run in the
build container where /root/reference exists:

    python tests/golden/make_golden.py

The fixture travels to the GPU box, the reference does not.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    R = oracle.ref()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20261018)
    out = {"polys": R.polys(), "edge_symm": R.edge_symm()}
    # encoder vectors: several lengths, incl. the SURVEY 8(c) KAT
    kat = np.array([0xA5, 0x3C, 0xFF, 0x01], dtype=np.uint8)
    out["kat_msg"], out["kat_segs"] = kat, R.encode(kat)
    for bits in (8, 64, 256, 2048, 4096):
        msgs = rng.integers(0, 256, (24, bits // 8), dtype=np.uint8)
        segs = R.encode_batch(msgs)
        out["msg_%d" % bits], out["segs_%d" % bits] = msgs, segs
        for tag, p in (("p0", 0.0), ("p02", 0.0226), ("p06", 0.0559), ("p50", 0.5)):
            flips = rng.random(segs.shape + (2,)) < p
            noisy = segs ^ (flips[..., 0].astype(np.uint8) | (flips[..., 1].astype(np.uint8) << 1))
            out["noisy_%d_%s" % (bits, tag)] = noisy
            out["dec_%d_%s" % (bits, tag)] = R.decode_batch(noisy, bits + 6)
    # streaming: metrics after every 64-segment call on a pure-noise 2048-bit frame
    noise = rng.integers(0, 4, 2054, dtype=np.uint8)
    dec, metrics = R.decode_chunked(noise, 64)
    out["stream_noise"], out["stream_dec"], out["stream_metrics"] = noise, dec, metrics
    # berTestK7 golden integers (berTestK7/berTestK7.c, srand(9865)), reduced and full
    counts = []
    for i, p in enumerate((5.585640e-02, 3.716174e-02, 2.262231e-02)):
        counts.append(R.bertest(9865 if i == 0 else 0, 10000, 256, p)[0])
    out["ber_counts_full"] = np.stack(counts)
    c, noisy, msgs = R.bertest(9865, 64, 256, 5.585640e-02, want_data=True)
    out["ber64_counts"], out["ber64_noisy"], out["ber64_msgs"] = c, noisy, msgs
    path = os.path.join(ROOT, "tests", "golden", "k7_reference_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
