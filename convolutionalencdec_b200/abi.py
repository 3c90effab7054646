"""ctypes binding of include/ced_abi.h (libced_cuda.so).

Device buffers are torch CUDA uint8 tensors; only their ``data_ptr()`` crosses
the ABI.  Every wrapper raises :class:`CedError` on a non-zero return code --
nothing here computes an encode or a decode on the host.
"""
import ctypes as C
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

_u8p = C.c_void_p


class CedError(RuntimeError):
    pass


class _CodeStruct(C.Structure):
    _fields_ = [("constraintLen", C.c_int32), ("codedBits", C.c_int32), ("gen", C.c_uint64 * 8)]


class Code:
    """Code parameters (Proakis convention, src/defaultParams/convCodeParams.c:6)."""

    def __init__(self, K, g):
        self.K, self.n, self.g = int(K), len(g), tuple(int(x) for x in g)
        self.S = self.K - 1
        self._c = _CodeStruct(self.K, self.n, (C.c_uint64 * 8)(*self.g))

    def segments(self, frame_bits):
        return frame_bits + self.S

    def __repr__(self):
        return "Code(K=%d, g=(%s))" % (self.K, ", ".join(oct(x) for x in self.g))


K7_DEFAULT = Code(7, (0o113, 0o171))    # the reference's defaultParams
K7_TEXTBOOK = Code(7, (0o133, 0o171))   # the MATLAB scripts' generators (scripts/matlab/viterbiBEREstimate.m:11)


def lib_path(name="libced_cuda.so"):
    return os.path.join(HERE, name)


def exported_abi_symbols():
    """Function names declared in include/ced_abi.h."""
    with open(os.path.join(ROOT, "include", "ced_abi.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ced_[a-z_0-9]+)\s*\(", text)))


_LIB = None


def load_abi():
    """Load libced_cuda.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise CedError("%s is missing: run `make cuda host` (or __graft_entry__.build()); "
                       "there is no CPU fallback" % path)
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    sz, i, vp, u64 = C.c_size_t, C.c_int, C.c_void_p, C.c_uint64
    codep = C.POINTER(_CodeStruct)
    lib.ced_device_count.restype = i
    lib.ced_last_error.restype = C.c_char_p
    lib.ced_ctx_create.argtypes = [i, C.POINTER(vp)]
    lib.ced_ctx_destroy.argtypes = [vp]
    lib.ced_ctx_destroy.restype = None
    lib.ced_ctx_device.argtypes = [vp]
    lib.ced_default_ctx.restype = vp
    lib.ced_sync.argtypes = [vp, vp]
    lib.ced_launch_count.argtypes = [vp]
    lib.ced_launch_count.restype = u64
    lib.ced_ctx_set_profiling.argtypes = [vp, i]
    lib.ced_ctx_last_kernel_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.ced_ctx_last_fallback_frames.argtypes = [vp, C.POINTER(i)]
    lib.ced_probe_int_peak.argtypes = [vp, i, C.POINTER(C.c_double)]
    lib.ced_decode_batch.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_encode_batch.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_decode_batch_packed.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_stream_server_stats.argtypes = [C.POINTER(u64), C.POINTER(u64)]
    lib.ced_encode_batch_k.argtypes = [vp, codep, i, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_decode_batch_k.argtypes = [vp, codep, i, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_decode_batch_packed_host.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz]
    lib.ced_pack_symbols.argtypes = [vp, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_host_pack_symbols.argtypes = [_u8p, sz, i, i, _u8p, sz, i]
    lib.ced_slice_soft_symbols.argtypes = [vp, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_decode_batch_soft.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_decode_batch_softq.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_decode_batch_softq_host.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz]
    lib.ced_quantize_soft.argtypes = [vp, _u8p, sz, i, i, C.c_double, _u8p, sz, vp]
    lib.ced_slice_soft_to_bytes.argtypes = [vp, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_awgn_channel.argtypes = [vp, _u8p, sz, i, i, _u8p, sz, C.c_double, C.c_double, u64, u64, vp, vp]
    lib.ced_shard_range.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
    lib.ced_shard_range.restype = None
    lib.ced_multi_create.argtypes = [C.POINTER(i), i, C.POINTER(vp)]
    lib.ced_multi_destroy.argtypes = [vp]
    lib.ced_multi_destroy.restype = None
    lib.ced_multi_device_count.argtypes = [vp]
    lib.ced_multi_ctx.argtypes = [vp, i]
    lib.ced_multi_ctx.restype = vp
    lib.ced_decode_batch_host_multi.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz]
    lib.ced_encode_batch_host_multi.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz]
    lib.ced_ber_allreduce.argtypes = [vp, C.POINTER(vp), i]
    lib.ced_nccl_version.restype = i
    lib.ced_probe_copy_ceiling.argtypes = [vp, sz, i, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ced_multi_probe_copy_ceiling.argtypes = [vp, sz, i, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    lib.ced_device_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    lib.ced_device_free.argtypes = [vp, vp]
    lib.ced_device_free.restype = None
    lib.ced_copy_to_device.argtypes = [vp, vp, vp, sz]
    lib.ced_copy_to_host.argtypes = [vp, vp, vp, sz]
    lib.ced_encode_batch_packed.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz, vp]
    lib.ced_decode_batch_host.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz]
    lib.ced_encode_batch_host.argtypes = [vp, codep, _u8p, sz, i, i, _u8p, sz]
    lib.ced_decode_scratch_bytes.argtypes = [i, i]
    lib.ced_decode_scratch_bytes.restype = sz
    lib.ced_host_register.argtypes = [vp, sz]
    lib.ced_host_unregister.argtypes = [vp]
    lib.ced_window_carry_bytes.argtypes = [i, i]
    lib.ced_window_carry_bytes.restype = sz
    lib.ced_decode_window_batch.argtypes = [vp, codep, _u8p, sz, i, i, u64, i, i, vp, _u8p, sz, vp]
    lib.ced_decode_window_batch_packed.argtypes = [vp, codep, _u8p, sz, i, i, u64, i, i, vp, _u8p, sz, vp]
    lib.ced_window_carry_bytes_code.restype = sz
    lib.ced_window_carry_bytes_code.argtypes = [codep, i, i]
    lib.ced_decode_window_batch_softq.argtypes = [vp, codep, _u8p, sz, i, i, u64, i, i, vp, _u8p, sz, vp]
    lib.ced_ber_count.argtypes = [vp, _u8p, sz, _u8p, sz, i, i, vp, vp]
    lib.ced_bsc_channel.argtypes = [vp, _u8p, sz, i, i, i, C.c_double, u64, u64, vp, vp]
    lib.ced_random_bytes.argtypes = [vp, _u8p, sz, i, i, u64, u64, vp]
    lib.ced_stream_surv_words.argtypes = [i]
    lib.ced_stream_decode.argtypes = [i, i, _u8p, _u8p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), vp,
                                      C.c_uint32, _u8p, i, _u8p, i]
    lib.ced_stream_encode.argtypes = [i, i, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), _u8p, i, _u8p, i]
    _LIB = lib
    return lib


def _check(lib, rc, what):
    if rc != 0:
        raise CedError("%s failed (%d): %s" % (what, rc, lib.ced_last_error().decode()))


def _stream_handle(stream):
    """cudaStream_t to launch on.  None means torch's CURRENT stream, so calls are ordered with
    the torch ops that produced the tensors.  Handle 0 (the legacy default stream) is passed as
    cudaStreamLegacy (1) because NULL means "the context's own stream" in ced_abi.h."""
    if stream is None:
        import torch
        handle = int(torch.cuda.current_stream().cuda_stream)
    else:
        handle = int(getattr(stream, "cuda_stream", stream))
    return C.c_void_p(handle if handle != 0 else 1)


class WindowDecoder:
    """nStreams continuous K=7 streams decoded a slice at a time; owns the carry block between calls."""

    def __init__(self, ctx, code, n_streams, depth, packed=False, softq=False):
        import torch
        self.ctx, self.code, self.n, self.depth, self.pos, self.packed = ctx, code, n_streams, depth, 0, packed
        self.softq = softq
        nbytes = ctx.lib.ced_window_carry_bytes_code(C.byref(code._c), n_streams, depth)
        if n_streams > 0 and nbytes == 0:
            raise ValueError("depth must be a multiple of 24, at least 24, and the code one the windowed decoder takes")
        self.carry = torch.empty(max(nbytes, 16), dtype=torch.uint8, device="cuda:%d" % ctx.device)

    def push(self, segs, last=False, out=None, stream=None, n_segments=None):
        """segs: CUDA uint8 [n_streams, >= slice length] (packed format: 4 segments per byte, pass n_segments);
        returns the decoded bytes that became final."""
        import torch
        n_seg = n_segments if n_segments is not None else segs.shape[1]
        if out is None:
            out = torch.empty((self.n, (n_seg + self.depth) // 8 + 1), dtype=torch.uint8, device=segs.device)
        fn = self.ctx.lib.ced_decode_window_batch_packed if self.packed else (
            self.ctx.lib.ced_decode_window_batch_softq if self.softq else self.ctx.lib.ced_decode_window_batch)
        rc = fn(self.ctx.h, C.byref(self.code._c), segs.data_ptr(), segs.stride(0),
                                                  self.n, n_seg, self.pos, self.depth, int(bool(last)),
                                                  self.carry.data_ptr(), out.data_ptr(), out.stride(0),
                                                  _stream_handle(stream))
        if rc < 0:
            _check(self.ctx.lib, rc, "ced_decode_window_batch")
        self.pos = 0 if last else self.pos + n_seg
        return out[:, :rc]


def shard_range(n_frames, n_shards, shard):
    """(first, count) of the contiguous frame range a shard owns (ced_shard_range; no GPU involved)."""
    first, count = C.c_int(0), C.c_int(0)
    load_abi().ced_shard_range(int(n_frames), int(n_shards), int(shard), C.byref(first), C.byref(count))
    return first.value, count.value


class MultiContext:
    """ced_multi: one process, several GPUs -- host batches sharded by the C library itself."""

    def __init__(self, devices=None):
        self.lib = load_abi()
        h = C.c_void_p()
        if devices:
            arr = (C.c_int * len(devices))(*devices)
            _check(self.lib, self.lib.ced_multi_create(arr, len(devices), C.byref(h)), "ced_multi_create")
        else:
            _check(self.lib, self.lib.ced_multi_create(None, 0, C.byref(h)), "ced_multi_create")
        self.h = h
        self.n_devices = int(self.lib.ced_multi_device_count(h))

    def ctx(self, i):
        """Borrowed Context of the i-th device (do not close it)."""
        c = Context.__new__(Context)
        c.lib, c.h, c.device, c.borrowed = self.lib, C.c_void_p(self.lib.ced_multi_ctx(self.h, i)), None, True
        c.device = int(self.lib.ced_ctx_device(c.h))
        return c

    def decode_batch_host(self, code, segs, frame_bits, out):
        sp, ss, sshape = Context._host(segs)
        op, os_, _ = Context._host(out)
        _check(self.lib, self.lib.ced_decode_batch_host_multi(self.h, C.byref(code._c), sp, ss, sshape[0], frame_bits,
                                                              op, os_), "ced_decode_batch_host_multi")
        return out

    def encode_batch_host(self, code, msgs, out):
        mp, ms, mshape = Context._host(msgs)
        op, os_, _ = Context._host(out)
        _check(self.lib, self.lib.ced_encode_batch_host_multi(self.h, C.byref(code._c), mp, ms, mshape[0], mshape[1],
                                                              op, os_), "ced_encode_batch_host_multi")
        return out

    def ber_allreduce(self, counters):
        """counters: one CUDA int64/uint64 tensor per device (same length), summed in place over NCCL."""
        ptrs = (C.c_void_p * len(counters))(*[c.data_ptr() for c in counters])
        _check(self.lib, self.lib.ced_ber_allreduce(self.h, ptrs, int(counters[0].numel())), "ced_ber_allreduce")

    def probe_copy_ceiling(self, bytes_per_device=256 << 20, reps=3):
        up, down = C.c_double(0), C.c_double(0)
        _check(self.lib, self.lib.ced_multi_probe_copy_ceiling(self.h, bytes_per_device, reps, C.byref(up), C.byref(down)),
               "ced_multi_probe_copy_ceiling")
        return up.value, down.value

    def close(self):
        if getattr(self, "h", None):
            self.lib.ced_multi_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Context:
    """One per GPU (ced_ctx): owns the compute/copy streams and survivor scratch."""

    def __init__(self, device=0):
        self.lib = load_abi()
        h = C.c_void_p()
        _check(self.lib, self.lib.ced_ctx_create(int(device), C.byref(h)), "ced_ctx_create")
        self.h, self.device = h, int(device)

    def close(self):
        if getattr(self, "h", None) and not getattr(self, "borrowed", False):
            self.lib.ced_ctx_destroy(self.h)
        self.h = None

    def probe_copy_ceiling(self, nbytes=256 << 20, reps=3):
        """(H2D, D2H) bytes per second of raw page-locked copies to / from this device."""
        up, down = C.c_double(0), C.c_double(0)
        _check(self.lib, self.lib.ced_probe_copy_ceiling(self.h, nbytes, reps, C.byref(up), C.byref(down)),
               "ced_probe_copy_ceiling")
        return up.value, down.value

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self, stream=None):
        _check(self.lib, self.lib.ced_sync(self.h, _stream_handle(stream)), "ced_sync")

    def set_profiling(self, enable=True):
        _check(self.lib, self.lib.ced_ctx_set_profiling(self.h, int(bool(enable))), "ced_ctx_set_profiling")

    def last_kernel_ms(self):
        """(forward_ms, traceback_ms) of the most recent decode_batch, CUDA-event timed."""
        ms = (C.c_float * 2)()
        _check(self.lib, self.lib.ced_ctx_last_kernel_ms(self.h, ms), "ced_ctx_last_kernel_ms")
        return float(ms[0]), float(ms[1])

    def last_fallback_frames(self):
        """frames of the most recent fused decode that were decoded again by the two-kernel path"""
        n = C.c_int(0)
        _check(self.lib, self.lib.ced_ctx_last_fallback_frames(self.h, C.byref(n)), "ced_ctx_last_fallback_frames")
        return n.value

    def probe_int_peak(self, mode=0):
        """32-bit lane-ops/s of a dependent-free LOP3 stream (mode 1: with IMAD co-issue)."""
        v = C.c_double(0)
        _check(self.lib, self.lib.ced_probe_int_peak(self.h, int(mode), C.byref(v)), "ced_probe_int_peak")
        return v.value

    @property
    def launches(self):
        return int(self.lib.ced_launch_count(self.h))

    # ---- device-resident batches (torch uint8 CUDA tensors, 2-D [frames, stride]) ----
    def decode_batch(self, code, segs, frame_bits, out=None, stream=None, n_frames=None):
        import torch
        nf = segs.shape[0] if n_frames is None else n_frames
        if out is None:
            out = torch.empty((nf, frame_bits // 8), dtype=torch.uint8, device=segs.device)
        _check(self.lib, self.lib.ced_decode_batch(self.h, C.byref(code._c), segs.data_ptr(), segs.stride(0), nf,
                                                   frame_bits, out.data_ptr(), out.stride(0), _stream_handle(stream)),
               "ced_decode_batch")
        return out

    def decode_batch_packed(self, code, packed, frame_bits, out=None, stream=None):
        import torch
        nf = packed.shape[0]
        if out is None:
            out = torch.empty((nf, frame_bits // 8), dtype=torch.uint8, device=packed.device)
        _check(self.lib, self.lib.ced_decode_batch_packed(self.h, C.byref(code._c), packed.data_ptr(),
                                                          packed.stride(0), nf, frame_bits, out.data_ptr(),
                                                          out.stride(0), _stream_handle(stream)),
               "ced_decode_batch_packed")
        return out

    def host_register(self, array):
        """Page-lock a numpy array the caller keeps passing to the *_host calls."""
        _check(self.lib, self.lib.ced_host_register(array.ctypes.data, array.nbytes), "ced_host_register")

    def host_unregister(self, array):
        _check(self.lib, self.lib.ced_host_unregister(array.ctypes.data), "ced_host_unregister")

    def window_decoder(self, code, n_streams, depth=48, packed=False, softq=False):
        """Continuous streams with windowed traceback (ced_decode_window_batch[_packed | _softq])."""
        return WindowDecoder(self, code, n_streams, depth, packed, softq)

    def pack_symbols(self, segs, segs_per_frame, out=None, stream=None, packed_stride=None):
        import torch
        nf = segs.shape[0]
        if out is None:
            out = torch.zeros((nf, packed_stride or (segs_per_frame + 3) // 4), dtype=torch.uint8, device=segs.device)
        _check(self.lib, self.lib.ced_pack_symbols(self.h, segs.data_ptr(), segs.stride(0), nf, segs_per_frame,
                                                   out.data_ptr(), out.stride(0), _stream_handle(stream)),
               "ced_pack_symbols")
        return out

    def slice_soft_symbols(self, soft, segs_per_frame, out=None, stream=None, packed_stride=None):
        """soft: int8 CUDA tensor [frames, >= 2*segs_per_frame] -> packed hard symbols."""
        import torch
        nf = soft.shape[0]
        if out is None:
            out = torch.zeros((nf, packed_stride or (segs_per_frame + 3) // 4), dtype=torch.uint8, device=soft.device)
        _check(self.lib, self.lib.ced_slice_soft_symbols(self.h, soft.data_ptr(), soft.stride(0), nf, segs_per_frame,
                                                         out.data_ptr(), out.stride(0), _stream_handle(stream)),
               "ced_slice_soft_symbols")
        return out

    def decode_batch_soft(self, code, soft, frame_bits, out=None, stream=None):
        """soft: int8 CUDA tensor [frames, >= 2*(frame_bits+6)], 16-byte aligned rows -> decoded bytes
        (true soft-decision decoding, ced_decode_batch_soft)."""
        import torch
        nf = soft.shape[0]
        if out is None:
            out = torch.empty((nf, frame_bits // 8), dtype=torch.uint8, device=soft.device)
        _check(self.lib, self.lib.ced_decode_batch_soft(self.h, C.byref(code._c), soft.data_ptr(), soft.stride(0), nf,
                                                        frame_bits, out.data_ptr(), out.stride(0),
                                                        _stream_handle(stream)), "ced_decode_batch_soft")
        return out

    def decode_batch_softq(self, code, syms, frame_bits, out=None, stream=None):
        """syms: uint8 CUDA tensor [frames, >= frame_bits+6], one byte per segment x0 | x1 << 3 (3-bit soft decisions)."""
        import torch
        nf = syms.shape[0]
        if out is None:
            out = torch.empty((nf, frame_bits // 8), dtype=torch.uint8, device=syms.device)
        _check(self.lib, self.lib.ced_decode_batch_softq(self.h, C.byref(code._c), syms.data_ptr(), syms.stride(0), nf,
                                                         frame_bits, out.data_ptr(), out.stride(0),
                                                         _stream_handle(stream)), "ced_decode_batch_softq")
        return out

    def decode_batch_softq_host(self, code, syms, frame_bits, out):
        """syms / out: host uint8 arrays (numpy or CPU tensors), 2-D; synchronous (ced_decode_batch_softq_host)."""
        sp, ss, sshape = Context._host(syms)
        op, os_, _ = Context._host(out)
        _check(self.lib, self.lib.ced_decode_batch_softq_host(self.h, C.byref(code._c), sp, ss, sshape[0], frame_bits, op, os_),
               "ced_decode_batch_softq_host")
        return out

    def quantize_soft(self, soft, segs_per_frame, delta, out=None, stream=None, sym_stride=None):
        """int8 reliabilities [frames, >= 2*segs] -> 3-bit soft symbols [frames, sym_stride] (ced_quantize_soft)."""
        import torch
        nf = soft.shape[0]
        if out is None:
            out = torch.zeros((nf, sym_stride or (segs_per_frame + 15) // 16 * 16), dtype=torch.uint8, device=soft.device)
        _check(self.lib, self.lib.ced_quantize_soft(self.h, soft.data_ptr(), soft.stride(0), nf, segs_per_frame,
                                                    float(delta), out.data_ptr(), out.stride(0), _stream_handle(stream)),
               "ced_quantize_soft")
        return out

    def slice_soft_to_bytes(self, soft, segs_per_frame, out=None, stream=None, seg_stride=None):
        import torch
        nf = soft.shape[0]
        if out is None:
            out = torch.zeros((nf, seg_stride or segs_per_frame), dtype=torch.uint8, device=soft.device)
        _check(self.lib, self.lib.ced_slice_soft_to_bytes(self.h, soft.data_ptr(), soft.stride(0), nf, segs_per_frame,
                                                          out.data_ptr(), out.stride(0), _stream_handle(stream)),
               "ced_slice_soft_to_bytes")
        return out

    def awgn_channel(self, segs, segs_per_frame, ebn0_db, seed, amplitude=32.0, first_frame=0, counters=None, out=None,
                     stream=None, soft_stride=None, rate=0.5):
        """BPSK + AWGN at Eb/N0 (dB) on byte-per-segment symbols -> int8 soft symbols [frames, soft_stride]."""
        import torch
        nf = segs.shape[0]
        if out is None:
            stride = soft_stride or (2 * segs_per_frame + 15) // 16 * 16
            out = torch.zeros((nf, stride), dtype=torch.int8, device=segs.device)
        sigma = (2.0 * rate * 10.0 ** (ebn0_db / 10.0)) ** -0.5
        _check(self.lib, self.lib.ced_awgn_channel(self.h, segs.data_ptr(), segs.stride(0), nf, segs_per_frame,
                                                   out.data_ptr(), out.stride(0), float(amplitude), float(sigma),
                                                   int(seed), int(first_frame),
                                                   counters.data_ptr() if counters is not None else None,
                                                   _stream_handle(stream)), "ced_awgn_channel")
        return out

    # ---- rate-k/n codes with k > 1: code.g are the k*K-bit generators, code.K the constraint length ----
    def encode_batch_k(self, code, k, msgs, out=None, stream=None, seg_stride=None):
        import torch
        nf, nb = msgs.shape
        T = 8 * nb // k + code.S
        if out is None:
            out = torch.zeros((nf, seg_stride or T), dtype=torch.uint8, device=msgs.device)
        _check(self.lib, self.lib.ced_encode_batch_k(self.h, C.byref(code._c), k, msgs.data_ptr(), msgs.stride(0), nf, nb,
                                                     out.data_ptr(), out.stride(0), _stream_handle(stream)),
               "ced_encode_batch_k")
        return out

    def decode_batch_k(self, code, k, segs, frame_bits, out=None, stream=None):
        import torch
        nf = segs.shape[0]
        if out is None:
            out = torch.empty((nf, frame_bits // 8), dtype=torch.uint8, device=segs.device)
        _check(self.lib, self.lib.ced_decode_batch_k(self.h, C.byref(code._c), k, segs.data_ptr(), segs.stride(0), nf,
                                                     frame_bits, out.data_ptr(), out.stride(0), _stream_handle(stream)),
               "ced_decode_batch_k")
        return out

    def encode_batch_packed(self, code, msgs, out=None, stream=None, packed_stride=None):
        import torch
        nf, nb = msgs.shape
        pb = (8 * nb + code.S + 3) // 4
        if out is None:
            out = torch.zeros((nf, packed_stride or pb), dtype=torch.uint8, device=msgs.device)
        _check(self.lib, self.lib.ced_encode_batch_packed(self.h, C.byref(code._c), msgs.data_ptr(), msgs.stride(0),
                                                          nf, nb, out.data_ptr(), out.stride(0),
                                                          _stream_handle(stream)), "ced_encode_batch_packed")
        return out

    def encode_batch(self, code, msgs, out=None, stream=None, seg_stride=None):
        import torch
        nf, nb = msgs.shape
        T = 8 * nb + code.S
        if out is None:
            out = torch.empty((nf, seg_stride or T), dtype=torch.uint8, device=msgs.device)
        _check(self.lib, self.lib.ced_encode_batch(self.h, C.byref(code._c), msgs.data_ptr(), msgs.stride(0), nf, nb,
                                                   out.data_ptr(), out.stride(0), _stream_handle(stream)),
               "ced_encode_batch")
        return out

    def ber_count(self, a, b, counters, stream=None):
        nf, nb = a.shape
        _check(self.lib, self.lib.ced_ber_count(self.h, a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), nf, nb,
                                                counters.data_ptr(), _stream_handle(stream)), "ced_ber_count")

    def bsc_channel(self, segs, segs_per_frame, n, p, seed, first_frame=0, counters=None, stream=None):
        _check(self.lib, self.lib.ced_bsc_channel(self.h, segs.data_ptr(), segs.stride(0), segs.shape[0],
                                                  segs_per_frame, n, float(p), int(seed), int(first_frame),
                                                  counters.data_ptr() if counters is not None else None,
                                                  _stream_handle(stream)), "ced_bsc_channel")

    def random_bytes(self, msgs, seed, first_frame=0, stream=None):
        _check(self.lib, self.lib.ced_random_bytes(self.h, msgs.data_ptr(), msgs.stride(0), msgs.shape[0],
                                                   msgs.shape[1], int(seed), int(first_frame),
                                                   _stream_handle(stream)), "ced_random_bytes")

    # ---- host buffers (numpy arrays or pinned torch CPU tensors) ----
    @staticmethod
    def _host(a):
        if hasattr(a, "data_ptr"):
            return a.data_ptr(), a.stride(0), a.shape
        return a.ctypes.data, a.strides[0], a.shape

    def decode_batch_host(self, code, segs, frame_bits, out):
        sp, ss, sshape = self._host(segs)
        op, os_, _ = self._host(out)
        _check(self.lib, self.lib.ced_decode_batch_host(self.h, C.byref(code._c), sp, ss, sshape[0], frame_bits, op,
                                                        os_), "ced_decode_batch_host")
        return out

    def decode_batch_packed_host(self, code, packed, frame_bits, out):
        sp, ss, sshape = self._host(packed)
        op, os_, _ = self._host(out)
        _check(self.lib, self.lib.ced_decode_batch_packed_host(self.h, C.byref(code._c), sp, ss, sshape[0],
                                                               frame_bits, op, os_), "ced_decode_batch_packed_host")
        return out

    def encode_batch_host(self, code, msgs, out):
        mp, ms, mshape = self._host(msgs)
        op, os_, _ = self._host(out)
        _check(self.lib, self.lib.ced_encode_batch_host(self.h, C.byref(code._c), mp, ms, mshape[0], mshape[1], op,
                                                        os_), "ced_encode_batch_host")
        return out
