"""ctypes view of the host-side C drop-in library (libconvencdec_k7.so / _k3.so)
using the reference's own names, so tests read like the reference's drivers:

    api = RefApi("k7")
    enc = api.encoder(); enc.resetConvEncoder(); enc.initConvEncoder()
    segs = enc.convEnc(msg, last=True)
    dec = api.decoder(); dec.VITERBI_RESET(); dec.VITERBI_INIT(); api.viterbiConfigCheck()
    out = dec.VITERBI_DECODER_HARD(segs, last=True)

Every convEnc / VITERBI_DECODER_HARD call goes host C -> ced_abi -> CUDA kernel.
"""
import ctypes as C
import os

import numpy as np

from .abi import CedError, HERE, load_abi

_u8p = C.POINTER(C.c_uint8)


def _p(a):
    return a.ctypes.data_as(_u8p)


class RefApi:
    def __init__(self, params="k7"):
        load_abi()  # libced_cuda.so first (RTLD_GLOBAL) so the host library resolves against it
        path = os.path.join(HERE, "libconvencdec_%s.so" % params)
        if not os.path.exists(path):
            raise CedError("%s is missing: run `make host`" % path)
        self.lib = lib = C.CDLL(path)
        for f in ("ced_sizeof_encoder_state", "ced_sizeof_decoder_state", "ced_offsetof_node_metrics_cur",
                  "ced_offsetof_edge_symm", "ced_offsetof_polynomials"):
            getattr(lib, f).restype = C.c_size_t
        lib.ced_param_num_states.restype = C.c_ulong
        self.K, self.n, self.N = lib.ced_param_K(), lib.ced_param_n(), int(lib.ced_param_num_states())
        self.S = self.K - 1
        lib.convEnc.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, C.c_bool]
        lib.convEnc.restype = C.c_int
        lib.convEncOneInput.argtypes = [C.c_void_p, C.c_uint8]
        lib.viterbiDecoderHardButterflyk1.argtypes = [C.c_void_p, _u8p, _u8p, C.c_int, C.c_bool]
        lib.viterbiDecoderHardButterflyk1.restype = C.c_int
        lib.calcHammingDist.argtypes = [C.c_uint8, C.c_uint8, C.c_int]
        lib.calcHammingDist.restype = C.c_uint8
        for f in ("resetConvEncoder", "initConvEncoder", "viterbiInitButterflyk1",
                  "resetViterbiDecoderHardButterflyk1"):
            getattr(lib, f).argtypes = [C.c_void_p]
            getattr(lib, f).restype = None
        self.g = [int(x) for x in (C.c_uint64 * self.n).in_dll(lib, "g")]

    def viterbiConfigCheck(self):
        return self.lib.viterbiConfigCheck()

    def calcHammingDist(self, a, b, bits):
        return int(self.lib.calcHammingDist(a, b, bits))

    def encoder(self):
        return _Encoder(self)

    def decoder(self):
        return _Decoder(self)


class _Encoder:
    def __init__(self, api):
        self.api, self.lib = api, api.lib
        self.buf = np.full(int(self.lib.ced_sizeof_encoder_state()) + 8, 0xA5, dtype=np.uint8)  # "stack garbage"
        self.p = self.buf.ctypes.data

    def resetConvEncoder(self):
        self.lib.resetConvEncoder(self.p)

    def initConvEncoder(self):
        self.lib.initConvEncoder(self.p)

    def polynomials(self):
        off = int(self.lib.ced_offsetof_polynomials())
        return self.buf[off:off + self.api.n].copy()

    def convEnc(self, uncoded, last):
        uncoded = np.ascontiguousarray(uncoded, dtype=np.uint8)
        segs = np.zeros(8 * uncoded.size + self.api.S, dtype=np.uint8)
        cnt = self.lib.convEnc(self.p, _p(uncoded), _p(segs), uncoded.size, bool(last))
        return segs[:cnt].copy()


class _Decoder:
    def __init__(self, api):
        self.api, self.lib = api, api.lib
        size = int(self.lib.ced_sizeof_decoder_state())
        raw = np.full(size + 64, 0x5A, dtype=np.uint8)  # uninitialised-stack stand-in, 64-byte aligned view
        shift = (-raw.ctypes.data) % 64
        self.buf = raw[shift:shift + size]
        self._raw = raw
        self.p = self.buf.ctypes.data

    def VITERBI_RESET(self):
        self.lib.resetViterbiDecoderHardButterflyk1(self.p)

    def VITERBI_INIT(self):
        self.lib.viterbiInitButterflyk1(self.p)

    def edgeCodedBitsSymm(self):
        off = int(self.lib.ced_offsetof_edge_symm())
        return self.buf[off:off + self.api.N // 2].copy()

    def nodeMetricsCur(self):
        """(*state.nodeMetricsCur)[i] as handTraced.c:72-111 reads it."""
        off = int(self.lib.ced_offsetof_node_metrics_cur())
        ptr = int(np.frombuffer(self.buf[off:off + 8].tobytes(), dtype=np.uint64)[0])
        return np.ctypeslib.as_array((C.c_uint8 * self.api.N).from_address(ptr)).copy()

    def VITERBI_DECODER_HARD(self, codedSegments, last, max_bytes=2049):
        codedSegments = np.ascontiguousarray(codedSegments, dtype=np.uint8)
        out = np.zeros(max_bytes, dtype=np.uint8)
        src = codedSegments if codedSegments.size else np.zeros(1, dtype=np.uint8)
        nb = self.lib.viterbiDecoderHardButterflyk1(self.p, _p(src), _p(out), codedSegments.size, bool(last))
        return out[:nb].copy()
