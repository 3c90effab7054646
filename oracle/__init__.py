"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the parity checker.

`port()` loads oracle/libced_oracle.so (the C restatement, ced_oracle.c) and
`ref()` loads oracle/_ref/libced_ref_k7_*.so (the unmodified reference sources
compiled by oracle/Makefile).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg may import this package; the
product package convolutionalencdec_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
K7_G = (0o113, 0o171)          # src/defaultParams/convCodeParams.c:6
K3_G = (0b111, 0b110)          # handTracedTest/testParams/convCodeParams.c:6

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)


def _p(a, t=_u8p):
    return a.ctypes.data_as(t)


def build(quiet=True):
    """(Re)build the oracle libraries; `ref` is a no-op without /root/reference."""
    subprocess.run(["make", "-C", HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            flags = f.read()
        return all(x in flags for x in ("avx512f", "avx512bw", "avx512vl"))
    except OSError:
        return False


class Port:
    """The C restatement, run-time parameterised by (K, n, g)."""

    def __init__(self):
        path = os.path.join(HERE, "libced_oracle.so")
        if not os.path.exists(path):
            build()
        self.lib = lib = C.CDLL(path)
        lib.orc_encode.restype = C.c_int
        lib.orc_encode.argtypes = [C.c_int, C.c_int, _u64p, C.POINTER(C.c_uint32), _u8p, C.c_int, _u8p, C.c_int]
        lib.orc_dec_new.restype = C.c_void_p
        lib.orc_dec_new.argtypes = [C.c_int, C.c_int, _u64p, C.c_int, C.c_int]
        lib.orc_dec_free.argtypes = [C.c_void_p]
        lib.orc_dec_reset.argtypes = [C.c_void_p]
        lib.orc_dec_metrics.argtypes = [C.c_void_p, _u8p]
        lib.orc_dec_edge_symm.argtypes = [C.c_void_p, _u8p]
        lib.orc_dec_edge.argtypes = [C.c_void_p, _u8p]
        lib.orc_dec_survivors.argtypes = [C.c_void_p, C.c_uint32, _u8p]
        lib.orc_dec_step.restype = C.c_int
        lib.orc_dec_step.argtypes = [C.c_void_p, _u8p, C.c_int, _u8p, C.c_int]
        lib.orc_decode_batch.restype = C.c_int
        lib.orc_decode_batch.argtypes = [C.c_int, C.c_int, _u64p, C.c_int, _u8p, C.c_size_t, C.c_int, C.c_int,
                                         _u8p, C.c_size_t]
        lib.orc_decode_window.restype = C.c_int
        lib.orc_decode_window.argtypes = [C.c_int, C.c_int, _u64p, _u8p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.orc_encode_batch.restype = C.c_int
        lib.orc_encode_batch.argtypes = [C.c_int, C.c_int, _u64p, _u8p, C.c_size_t, C.c_int, C.c_int, _u8p,
                                         C.c_size_t]
        lib.orc_taps.argtypes = [C.c_int, C.c_int, _u64p, C.POINTER(C.c_uint32)]
        lib.orc_srand.argtypes = [C.c_uint]
        lib.orc_bertest.restype = C.c_int
        lib.orc_bertest.argtypes = [C.c_int, C.c_int, _u64p, C.c_int, C.c_int, C.c_double, _i64p, _u8p, _u8p]
        lib.orc_speed_decode.restype = C.c_int64
        lib.orc_speed_decode.argtypes = [C.c_int, C.c_int, _u64p, _u8p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                         C.c_double, C.POINTER(C.c_double)]
        lib.orc_hamming.restype = C.c_uint8
        lib.orc_hamming.argtypes = [C.c_uint8, C.c_uint8, C.c_int]
        lib.orc_window_ber.restype = C.c_int
        lib.orc_window_ber.argtypes = [C.c_int, C.c_int, _u64p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                       C.c_uint64, C.c_int, _i64p]
        lib.orc_decode_soft_batch.restype = C.c_int
        lib.orc_decode_soft_batch.argtypes = [C.c_int, C.c_int, _u64p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                              _u8p, C.c_size_t]

        lib.orc_decode_window2.restype = C.c_int
        lib.orc_decode_window2.argtypes = [C.c_int, C.c_int, _u64p, C.c_int, _u8p, C.c_int, C.c_int, C.c_int, _u8p]
        lib.orc_decode_window_soft.restype = C.c_int
        lib.orc_decode_window_soft.argtypes = [C.c_int, C.c_int, _u64p, C.c_void_p, C.c_int, C.c_int, C.c_int, _u8p]
        _u32p = C.POINTER(C.c_uint32)
        lib.orck_encode_batch.restype = C.c_int
        lib.orck_encode_batch.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _u8p, C.c_size_t, C.c_int, C.c_int, _u8p,
                                          C.c_size_t]
        lib.orck_decode_batch.restype = C.c_int
        lib.orck_decode_batch.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _u8p, C.c_size_t, C.c_int, C.c_int, _u8p,
                                          C.c_size_t]
        lib.orck_edges.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _u8p]
        lib.orck_metrics.restype = C.c_int
        lib.orck_metrics.argtypes = [C.c_int, C.c_int, C.c_int, _u64p, _u8p, C.c_int, _u32p]

    @staticmethod
    def _g(g):
        return (C.c_uint64 * len(g))(*g)

    # ---- k > 1 (ced_oracle_k.c): one k*K-bit shift register, k bits per segment ----
    def encode_batch_k(self, K, k, g, msgs, seg_stride=None):
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        nf, nb = msgs.shape
        T = 8 * nb // k + K - 1
        stride = seg_stride or T
        segs = np.zeros((nf, stride), dtype=np.uint8)
        assert self.lib.orck_encode_batch(K, k, len(g), self._g(g), _p(msgs), nb, nf, nb, _p(segs), stride) == 0
        return segs

    def decode_batch_k(self, K, k, g, segs, T):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        nf, stride = segs.shape
        nbytes = (T - (K - 1)) * k // 8
        out = np.zeros((nf, nbytes), dtype=np.uint8)
        assert self.lib.orck_decode_batch(K, k, len(g), self._g(g), _p(segs), stride, nf, T, _p(out), nbytes) == 0
        return out

    def edges_k(self, K, k, g):
        N = 1 << (k * (K - 1))
        out = np.zeros((1 << k, N), dtype=np.uint8)
        self.lib.orck_edges(K, k, len(g), self._g(g), _p(out))
        return out

    def metrics_k(self, K, k, g, segs):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        N = 1 << (k * (K - 1))
        out = np.zeros((segs.size, N), dtype=np.uint32)
        assert self.lib.orck_metrics(K, k, len(g), self._g(g), _p(segs), segs.size,
                                     out.ctypes.data_as(C.POINTER(C.c_uint32))) == 0
        return out

    def taps(self, K, g):
        out = (C.c_uint32 * len(g))()
        self.lib.orc_taps(K, len(g), self._g(g), out)
        return list(out)

    def encode(self, K, g, msg, last=True, reg=0):
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        segs = np.zeros(8 * msg.size + (K - 1), dtype=np.uint8)
        r = C.c_uint32(reg)
        cnt = self.lib.orc_encode(K, len(g), self._g(g), C.byref(r), _p(msg), msg.size, _p(segs), int(last))
        return segs[:cnt].copy(), r.value

    def encode_batch(self, K, g, msgs, seg_stride=None):
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        nf, nb = msgs.shape
        T = 8 * nb + K - 1
        stride = seg_stride or T
        segs = np.zeros((nf, stride), dtype=np.uint8)
        self.lib.orc_encode_batch(K, len(g), self._g(g), _p(msgs), nb, nf, nb, _p(segs), stride)
        return segs

    def decode_batch(self, K, g, segs, T, symmetric=True):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        nf, stride = segs.shape
        nbytes = (T - (K - 1) - 1) // 8 + 1
        out = np.zeros((nf, nbytes), dtype=np.uint8)
        rc = self.lib.orc_decode_batch(K, len(g), self._g(g), int(symmetric), _p(segs), stride, nf, T, _p(out),
                                       nbytes)
        assert rc == 0
        return out

    def window_ber(self, K, g, pkts, pkt_bytes, p, call_segs, depth, seed=1, threads=None):
        """BSC Monte-Carlo of orc_decode_window: {channel flips, coded bits, decoded errors, decoded bits}."""
        counts = np.zeros(4, dtype=np.int64)
        self.lib.orc_window_ber(K, len(g), self._g(g), pkts, pkt_bytes, float(p), call_segs, depth, seed,
                                threads or os.cpu_count() or 1, _p(counts, _i64p))
        return counts

    def decode_soft_batch(self, K, g, soft, T):
        """soft: int8 [frames, >= n*T], n values per segment (generator 0 first); semantics in
        ced_oracle.c:orc_dec_step_soft (reliability-weighted calcHammingDist, otherwise orc_dec_step)."""
        soft = np.ascontiguousarray(soft, dtype=np.int8)
        nf, stride = soft.shape
        nbytes = (T - (K - 1) - 1) // 8 + 1
        out = np.zeros((nf, nbytes), dtype=np.uint8)
        rc = self.lib.orc_decode_soft_batch(K, len(g), self._g(g), soft.ctypes.data, stride, nf, T, _p(out), nbytes)
        assert rc == 0
        return out

    def decode_window_soft(self, K, g, soft, call_segs, depth):
        """soft: int8 [n * total segments] of ONE stream; the windowing of decode_window around the soft recursion."""
        soft = np.ascontiguousarray(soft, dtype=np.int8)
        total = soft.size // len(g)
        out = np.zeros((total - (K - 1) + 7) // 8, dtype=np.uint8)
        rc = self.lib.orc_decode_window_soft(K, len(g), self._g(g), soft.ctypes.data, total, call_segs, depth, _p(out))
        assert rc == total - (K - 1)
        return out

    def decode_window(self, K, g, segs, call_segs, depth, symmetric=True):
        """One terminated stream decoded with windowed traceback (orc_decode_window: semantics defined there);
        symmetric=False: the general branch costs, for generators that do not tap both ends."""
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        T = segs.size
        out = np.zeros((T - (K - 1) + 7) // 8, dtype=np.uint8)
        rc = self.lib.orc_decode_window2(K, len(g), self._g(g), int(symmetric), _p(segs), T, int(call_segs), int(depth),
                                         _p(out))
        assert rc == T - (K - 1), rc
        return out

    def decoder(self, K, g, symmetric=True, max_segments=16390):
        return PortDecoder(self, K, g, symmetric, max_segments)

    def bertest(self, K, g, pkts, pkt_bytes, p, want_data=False):
        counts = np.zeros(4, dtype=np.int64)
        T = 8 * pkt_bytes + K - 1
        noisy = np.zeros((pkts, T), dtype=np.uint8) if want_data else None
        msgs = np.zeros((pkts, pkt_bytes), dtype=np.uint8) if want_data else None
        rc = self.lib.orc_bertest(K, len(g), self._g(g), pkts, pkt_bytes, float(p), _p(counts, _i64p),
                                  _p(noisy) if want_data else None, _p(msgs) if want_data else None)
        assert rc == 0
        return counts, noisy, msgs

    def speed_decode(self, K, g, segs, T, threads, seconds):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        nf, stride = segs.shape
        el = C.c_double(0)
        bits = self.lib.orc_speed_decode(K, len(g), self._g(g), _p(segs), stride, nf, T, threads, float(seconds),
                                         C.byref(el))
        return bits, el.value


class PortDecoder:
    def __init__(self, port, K, g, symmetric, max_segments):
        self.port, self.K, self.N = port, K, 1 << (K - 1)
        self.h = port.lib.orc_dec_new(K, len(g), Port._g(g), int(symmetric), max_segments)
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.port.lib.orc_dec_free(self.h)
            self.h = None

    def reset(self):
        self.port.lib.orc_dec_reset(self.h)

    def step(self, segs, last, out_bytes=2049):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        out = np.zeros(out_bytes, dtype=np.uint8)
        nb = self.port.lib.orc_dec_step(self.h, _p(segs), segs.size, _p(out), int(last))
        assert nb >= 0
        return out[:nb].copy()

    def metrics(self):
        m = np.zeros(self.N, dtype=np.uint8)
        self.port.lib.orc_dec_metrics(self.h, _p(m))
        return m

    def edge_symm(self):
        m = np.zeros(self.N // 2, dtype=np.uint8)
        self.port.lib.orc_dec_edge_symm(self.h, _p(m))
        return m

    def edge(self):
        m = np.zeros(2 * self.N, dtype=np.uint8)
        self.port.lib.orc_dec_edge(self.h, _p(m))
        return m.reshape(2, self.N)

    def survivors(self, t):
        m = np.zeros(self.N, dtype=np.uint8)
        self.port.lib.orc_dec_survivors(self.h, t, _p(m))
        return m


class Ref:
    """The unmodified reference (K=7 defaultParams) behind ref_harness.c."""

    def __init__(self):
        name = "libced_ref_k7_v4.so" if _cpu_has_avx512() else "libced_ref_k7_v3.so"
        path = os.path.join(HERE, "_ref", name)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.isa = name[-5:-3]
        self.lib = lib = C.CDLL(path)
        lib.refh_state_bytes.restype = C.c_size_t
        lib.refh_g.restype = C.c_uint64
        lib.refh_g.argtypes = [C.c_int]
        lib.refh_polys.argtypes = [_u8p]
        lib.refh_edge_symm.argtypes = [_u8p]
        lib.refh_encode.restype = C.c_int
        lib.refh_encode.argtypes = [_u8p, C.c_int, _u8p]
        lib.refh_encode_chunked.restype = C.c_int
        lib.refh_encode_chunked.argtypes = [_u8p, C.c_int, _u8p, C.c_int]
        lib.refh_decode_batch.restype = C.c_int
        lib.refh_decode_batch.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, _u8p, C.c_size_t]
        lib.refh_decode_batch_mt.restype = C.c_int
        lib.refh_decode_batch_mt.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, _u8p, C.c_size_t, C.c_int]
        lib.refh_decode_chunked.restype = C.c_int
        lib.refh_decode_chunked.argtypes = [_u8p, C.c_int, C.c_int, _u8p, _u8p]
        lib.refh_bertest.restype = C.c_int
        lib.refh_bertest.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_double, _i64p, _u8p, _u8p]
        lib.refh_speed_decode.restype = C.c_int64
        lib.refh_speed_decode.argtypes = [_u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, C.c_double,
                                          C.POINTER(C.c_double)]
        lib.refh_speed_encode.restype = C.c_int64
        lib.refh_speed_encode.argtypes = [_u8p, C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_double)]

    def polys(self):
        out = np.zeros(2, dtype=np.uint8)
        self.lib.refh_polys(_p(out))
        return out

    def edge_symm(self):
        out = np.zeros(32, dtype=np.uint8)
        self.lib.refh_edge_symm(_p(out))
        return out

    def encode(self, msg, chunk=None):
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        segs = np.zeros(8 * msg.size + 6, dtype=np.uint8)
        if chunk:
            cnt = self.lib.refh_encode_chunked(_p(msg), msg.size, _p(segs), chunk)
        else:
            cnt = self.lib.refh_encode(_p(msg), msg.size, _p(segs))
        return segs[:cnt].copy()

    def encode_batch(self, msgs):
        return np.stack([self.encode(m) for m in msgs])

    def decode_batch(self, segs, T):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        nf, stride = segs.shape
        nbytes = (T - 7) // 8 + 1
        out = np.zeros((nf, nbytes), dtype=np.uint8)
        self.lib.refh_decode_batch(_p(segs), stride, nf, T, _p(out), nbytes)
        return out

    def decode_batch_mt(self, segs, T, threads=None):
        """Every frame through the reference decoder, frames split over host threads."""
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        nf, stride = segs.shape
        nbytes = (T - 7) // 8 + 1
        out = np.zeros((nf, nbytes), dtype=np.uint8)
        self.lib.refh_decode_batch_mt(_p(segs), stride, nf, T, _p(out), nbytes, threads or (os.cpu_count() or 1))
        return out

    def decode_chunked(self, segs, chunk):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        T = segs.size
        calls = (T + chunk - 1) // chunk
        out = np.zeros((T - 7) // 8 + 1, dtype=np.uint8)
        metrics = np.zeros((calls, 64), dtype=np.uint8)
        nb = self.lib.refh_decode_chunked(_p(segs), T, chunk, _p(out), _p(metrics))
        return out[:nb].copy(), metrics

    def bertest(self, seed, pkts, pkt_bytes, p, want_data=False):
        counts = np.zeros(4, dtype=np.int64)
        T = 8 * pkt_bytes + 6
        noisy = np.zeros((pkts, T), dtype=np.uint8) if want_data else None
        msgs = np.zeros((pkts, pkt_bytes), dtype=np.uint8) if want_data else None
        self.lib.refh_bertest(seed, pkts, pkt_bytes, float(p), _p(counts, _i64p),
                              _p(noisy) if want_data else None, _p(msgs) if want_data else None)
        return counts, noisy, msgs

    def speed_decode(self, segs, T, threads, seconds):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        nf, stride = segs.shape
        el = C.c_double(0)
        bits = self.lib.refh_speed_decode(_p(segs), stride, nf, T, threads, float(seconds), C.byref(el))
        return bits, el.value

    def speed_encode(self, msgs, threads, seconds):
        msgs = np.ascontiguousarray(msgs, dtype=np.uint8)
        nf, nb = msgs.shape
        el = C.c_double(0)
        bits = self.lib.refh_speed_encode(_p(msgs), nf, nb, threads, float(seconds), C.byref(el))
        return bits, el.value


class RefK:
    """The unmodified reference built with k = 2 parameters (oracle/params/<name>/) behind ref_harness_k.c: encoder,
    edge labels and per-step path metrics of its generic decoder (whose traceback does not run at HEAD)."""

    def __init__(self, name):
        path = os.path.join(HERE, "_ref", "libced_refk_%s.so" % name)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = lib = C.CDLL(path)
        lib.refk_g.restype = C.c_uint64
        lib.refk_g.argtypes = [C.c_int]
        lib.refk_encode.restype = C.c_int
        lib.refk_encode.argtypes = [_u8p, C.c_int, _u8p]
        lib.refk_edges.argtypes = [_u8p]
        lib.refk_metrics.argtypes = [_u8p, C.c_int, C.POINTER(C.c_uint32)]
        self.K, self.k, self.n, self.N = lib.refk_K(), lib.refk_k(), lib.refk_n(), lib.refk_states()
        self.g = [int(lib.refk_g(i)) for i in range(self.n)]

    def encode(self, msg):
        msg = np.ascontiguousarray(msg, dtype=np.uint8)
        segs = np.zeros(8 * msg.size // self.k + self.K + 8, dtype=np.uint8)
        cnt = self.lib.refk_encode(_p(msg), msg.size, _p(segs))
        return segs[:cnt].copy()

    def edges(self):
        out = np.zeros((1 << self.k, self.N), dtype=np.uint8)
        self.lib.refk_edges(_p(out))
        return out

    def metrics(self, segs):
        segs = np.ascontiguousarray(segs, dtype=np.uint8)
        out = np.zeros((segs.size, self.N), dtype=np.uint32)
        self.lib.refk_metrics(_p(segs), segs.size, out.ctypes.data_as(C.POINTER(C.c_uint32)))
        return out


def refk(name):
    """Reference built with the k = 2 parameters `name`, or None when oracle/_ref was never built."""
    try:
        return RefK(name)
    except (FileNotFoundError, OSError, subprocess.CalledProcessError):
        return None


_PORT = None
_REF = None


def port():
    global _PORT
    if _PORT is None:
        _PORT = Port()
    return _PORT


def ref():
    """Returns the reference binding, or None when oracle/_ref was never built."""
    global _REF
    if _REF is None:
        try:
            _REF = Ref()
        except (FileNotFoundError, OSError, subprocess.CalledProcessError):
            _REF = False
    return _REF or None
