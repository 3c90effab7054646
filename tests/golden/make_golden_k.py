"""Golden vectors for the k = 2 codes (tests/golden/k2_vectors.npz).  Run in the build container, where
/root/reference exists: the encoder output is taken from the UNMODIFIED reference built with the k = 2 parameter
headers (oracle/_ref/libced_refk_*.so); the decoded bytes come from the restatement oracle/ced_oracle_k.c, whose
add-compare-select is checked here, step by step, against the reference's generic decoder before anything is written
(the reference's own k > 1 traceback does not run at HEAD).   python tests/golden/make_golden_k.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

CODES = {"k2K3n3": (3, 2, (0o27, 0o75, 0o72)), "k2K4n3": (4, 2, (0o236, 0o155, 0o337))}
P = oracle.port()
rng = np.random.default_rng(20261019)
out = {}
for name, (K, k, g) in CODES.items():
    R = oracle.refk(name)
    assert R is not None, "build oracle/_ref first (make oracle)"
    msgs = rng.integers(0, 256, (24, 64), dtype=np.uint8)
    segs = np.stack([R.encode(m) for m in msgs])
    assert np.array_equal(segs, P.encode_batch_k(K, k, g, msgs))
    noisy = segs.copy()
    for f in range(24):
        p = (0.0, 0.02, 0.08, 0.5)[f % 4]
        flips = rng.random(noisy[f].shape + (len(g),)) < p
        for j in range(len(g)):
            noisy[f] ^= (flips[..., j].astype(np.uint8) << j)
        assert np.array_equal(P.metrics_k(K, k, g, noisy[f]), R.metrics(noisy[f]))
    out[name + "_msgs"], out[name + "_segs"], out[name + "_noisy"] = msgs, segs, noisy
    out[name + "_decoded"] = P.decode_batch_k(K, k, g, noisy, noisy.shape[1])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "k2_vectors.npz"), **out)
print("wrote", sorted(out))
