"""The SIMD-in-word ACS algebra of csrc/trellis_swar.cuh (rotating state labels,
lane-phase compare constants, packed survivor layout, traceback bookkeeping) run
on the host for single frames and compared with the oracle -- decoded bytes and
every per-state decision."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from conftest import ROOT, bsc

u8p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)


@pytest.fixture(scope="module")
def sim():
    path = os.path.join(ROOT, "tests", "hostsim", "libswar_sim.so")
    subprocess.run(["make", "-C", ROOT, "hostsim"], check=True, stdout=subprocess.DEVNULL)
    lib = C.CDLL(path)
    lib.swar_sim_decode.argtypes = [u8p, C.c_int, u8p, u32p, u8p, C.c_int]
    lib.swar_sim_decision.argtypes = [u32p, C.c_int, C.c_int]
    return lib


def run(sim, row, period=96):
    T = row.size
    out = np.zeros((T - 6) // 8, dtype=np.uint8)
    surv = np.zeros(2 * T, dtype=np.uint32)
    mx = C.c_uint8(0)
    row = np.ascontiguousarray(row)
    sim.swar_sim_decode(row.ctypes.data_as(u8p), T, out.ctypes.data_as(u8p), surv.ctypes.data_as(u32p), C.byref(mx),
                        period)
    return out, surv, mx.value


@pytest.mark.parametrize("bits", [8, 16, 24, 48, 96, 512, 4096])
def test_decoded_bytes_match_oracle(sim, port, bits):
    rng = np.random.default_rng(bits)
    msgs = rng.integers(0, 256, (6, bits // 8), dtype=np.uint8)
    segs = port.encode_batch(7, oracle.K7_G, msgs)
    for p in (0.0, 0.04, 0.12, 0.5):
        noisy = bsc(rng, segs, p, junk_upper_bits=True)
        want = port.decode_batch(7, oracle.K7_G, noisy, bits + 6)
        for f in range(noisy.shape[0]):
            got, _, mx = run(sim, noisy[f])
            assert np.array_equal(got, want[f])
            assert mx < 126  # candidates must stay below 128 for the guard-bit compare


def test_every_decision_matches_oracle(sim, port):
    rng = np.random.default_rng(77)
    for p in (0.0, 0.1, 0.5):
        row = bsc(rng, port.encode_batch(7, oracle.K7_G, rng.integers(0, 256, (1, 32), dtype=np.uint8)), p)[0]
        T = row.size
        _, surv, _ = run(sim, row)
        d = port.decoder(7, oracle.K7_G)
        d.step(row, False)
        sp = surv.ctypes.data_as(u32p)
        for t in range(T):
            got = [sim.swar_sim_decision(sp, t, s) for s in range(64)]
            assert got == d.survivors(t).tolist(), t


def test_renorm_period_does_not_change_decisions(sim, port):
    rng = np.random.default_rng(3)
    row = rng.integers(0, 4, 1030, dtype=np.uint8)
    base, _, _ = run(sim, row, 96)
    for period in (6, 24, 48, 102):
        got, _, mx = run(sim, row, period)
        assert np.array_equal(got, base) and mx < 126



@pytest.mark.parametrize("total,call,depth,p", [(96 * 20 + 6, 96, 24, 0.08), (96 * 20 + 14, 192, 48, 0.06),
                                                (96 * 9 + 38, 288, 96, 0.10), (96 * 12 + 6, 96, 192, 0.12),
                                                (96 * 3 + 6, 96 * 8, 48, 0.05)])
def test_window_procedure_matches_oracle_definition(sim, port, total, call, depth, p):
    """Best-position start (bestPositionB), survivor-bit-index traceback and the slice bookkeeping of the windowed
    decoder, on the host, against orc_decode_window -- incl. high noise so that metric ties occur."""
    sim.swar_sim_window.argtypes = [u8p, C.c_int, C.c_int, C.c_int, u8p]
    rng = np.random.default_rng(total + depth)
    for _ in range(6):
        msg = rng.integers(0, 256, (1, (total - 6) // 8), dtype=np.uint8)
        noisy = np.ascontiguousarray(bsc(rng, port.encode_batch(7, oracle.K7_G, msg), p)[0])
        out = np.zeros((total - 6) // 8, dtype=np.uint8)
        rc = sim.swar_sim_window(noisy.ctypes.data_as(u8p), total, call, depth, out.ctypes.data_as(u8p))
        assert rc == total - 6
        assert np.array_equal(out, port.decode_window(7, oracle.K7_G, noisy, call, depth))


@pytest.mark.parametrize("g", [(0o113, 0o171), (0o133, 0o171), (0o171, 0o133), (0o117, 0o155), (0o135, 0o163),
                               (0o101, 0o177), (0o145, 0o175), (0o133, 0o171, 0o165), (0o133, 0o145, 0o175),
                               (0o175, 0o133, 0o171), (0o101, 0o101, 0o177)])
def test_runtime_table_path_matches_oracle(sim, port, g):
    """acsStepTable + buildStepTable (generators known only at run time, n = 2 or 3) against the oracle, per code;
    the largest metric must leave the guard bit free (renorm every 96 / 24 steps)."""
    sim.swar_sim_decode_rt.argtypes = [C.c_int, u32p, u8p, C.c_int, u8p]
    rng = np.random.default_rng(sum(g))
    n = len(g)
    gens = np.array(g, dtype=np.uint32)
    for bits, p in ((48, 0.0), (512, 0.03), (2048, 0.07), (1024, 0.5), (4096, 0.5)):
        msg = rng.integers(0, 256, (1, bits // 8), dtype=np.uint8)
        clean = port.encode_batch(7, list(g), msg)
        flips = rng.random(clean.shape + (n,)) < p
        noisy = clean.copy()
        for j in range(n):
            noisy ^= (flips[..., j].astype(np.uint8) << j)
        noisy = np.ascontiguousarray(noisy[0])
        out = np.zeros(bits // 8, dtype=np.uint8)
        mx = sim.swar_sim_decode_rt(n, gens.ctypes.data_as(u32p), noisy.ctypes.data_as(u8p), bits + 6,
                                    out.ctypes.data_as(u8p))
        assert mx < 128 - n, (g, bits, p, mx)
        assert np.array_equal(out, port.decode_batch(7, list(g), noisy[None, :], bits + 6)[0]), (g, bits, p)


GEN_CODES = [(3, (0b111, 0b110)), (3, (0b111, 0b101)), (3, (0b101, 0b011)), (3, (0b111, 0b101, 0b011)),
             (4, (0o15, 0o17)), (4, (0o13, 0o15, 0o17)), (4, (0o14, 0o07)),
             (5, (0o23, 0o35)), (5, (0o25, 0o33, 0o37)), (5, (0o30, 0o07)),
             (7, (0o133, 0o170)), (7, (0o066, 0o171)), (7, (0o113, 0o171)), (7, (0o133, 0o145, 0o174)),
             (9, (0o561, 0o753)), (9, (0o557, 0o663, 0o711)), (9, (0o460, 0o353)),
             (6, (0o53, 0o75)), (6, (0o47, 0o53, 0o75)), (6, (0o60, 0o17)), (8, (0o247, 0o371)), (8, (0o225, 0o331, 0o367)),
             (8, (0o300, 0o073))]


@pytest.mark.parametrize("K,g", GEN_CODES)
def test_generic_table_path_matches_oracle(sim, port, K, g):
    """genStep / buildGenTable / genTracebackStep (swar_generic.cuh: any k = 1 code with 4..256 states, generators
    that need not tap both ends) against the oracle's general butterflies; the largest metric must leave the guard
    bit free."""
    sim.swar_sim_decode_gen.argtypes = [C.c_int, C.c_int, u32p, u8p, C.c_int, u8p]
    rng = np.random.default_rng(K * 1000 + sum(g))
    n, S = len(g), K - 1
    gens = np.array(g, dtype=np.uint32)
    for bits, p in ((8, 0.0), (48, 0.0), (40, 0.2), (512, 0.03), (2048, 0.07), (1024, 0.5), (4096, 0.5)):
        msg = rng.integers(0, 256, (1, bits // 8), dtype=np.uint8)
        clean = port.encode_batch(K, list(g), msg)
        flips = rng.random(clean.shape + (n,)) < p
        noisy = clean.copy()
        for j in range(n):
            noisy ^= (flips[..., j].astype(np.uint8) << j)
        noisy |= rng.integers(0, 32, noisy.shape, dtype=np.uint8) << 3  # bits above n are ignored
        noisy = np.ascontiguousarray(noisy[0])
        out = np.zeros(bits // 8, dtype=np.uint8)
        mx = sim.swar_sim_decode_gen(K, n, gens.ctypes.data_as(u32p), noisy.ctypes.data_as(u8p), bits + S,
                                     out.ctypes.data_as(u8p))
        assert 0 <= mx < 128 - n, (K, g, bits, p, mx)
        want = port.decode_batch(K, list(g), noisy[None, :], bits + S, symmetric=False)[0]
        assert np.array_equal(out, want), (K, g, bits, p)


R4_CODES = [(2, (0o17, 0o06, 0o15)), (3, (0o27, 0o75, 0o72)), (4, (0o236, 0o155, 0o337)), (5, (0o1236, 0o0155, 0o1337)),
            (3, (0o53, 0o75)), (2, (0o13, 0o07, 0o15)), (4, (0o321, 0o256, 0o177))]


@pytest.mark.parametrize("K,g", R4_CODES)
def test_radix4_step_functions_match_the_k2_restatement(sim, port, K, g):
    """r4Step / buildR4Table / r4TracebackStep (swar_radix4.cuh, k = 2: four branches into every state, lowest edgeIn
    keeps a tie) against oracle/ced_oracle_k.c, which is pinned to the reference's own add-compare-select; the largest
    metric must leave the guard bit free."""
    sim.swar_sim_decode_r4.argtypes = [C.c_int, C.c_int, u32p, u8p, C.c_int, u8p]
    rng = np.random.default_rng(K * 100 + sum(g))
    n, S = len(g), K - 1
    gens = np.array(g, dtype=np.uint32)
    for nbytes, p in ((1, 0.0), (6, 0.0), (5, 0.2), (64, 0.03), (256, 0.07), (128, 0.5), (512, 0.5)):
        msg = rng.integers(0, 256, (1, nbytes), dtype=np.uint8)
        clean = port.encode_batch_k(K, 2, g, msg)
        T = clean.shape[1]
        flips = rng.random(clean.shape + (n,)) < p
        noisy = clean.copy()
        for j in range(n):
            noisy ^= (flips[..., j].astype(np.uint8) << j)
        want = port.decode_batch_k(K, 2, g, noisy, T)[0]
        junk = noisy | (rng.integers(0, 16, noisy.shape, dtype=np.uint8) << 4)    # bits above n are ignored
        row = np.ascontiguousarray(junk[0])
        out = np.zeros(nbytes, dtype=np.uint8)
        mx = sim.swar_sim_decode_r4(K, n, gens.ctypes.data_as(u32p), row.ctypes.data_as(u8p), T, out.ctypes.data_as(u8p))
        assert 0 <= mx < 128 - n, (K, g, nbytes, p, mx)
        assert np.array_equal(out, want), (K, g, nbytes, p)
