"""convolutionalencdec_b200 -- B200-native K=7 r=1/2 convolutional encoder and
hard-decision Viterbi decoder behind the ucb-cyarp/ConvolutionalEncDec C API.

The product is two C libraries built in-tree by the top-level Makefile:

* ``libced_cuda.so``      -- hand-written sm_100a kernels behind the ``extern "C"``
                             ABI of ``include/ced_abi.h``;
* ``libconvencdec_k7.so`` -- the host-side C drop-in for the reference's
  (and ``_k3`` for the      ``convEncode.h`` / ``viterbiDecoder.h`` /
  hand-traced test code)    ``viterbiDecoderButterflyk1.h``.

This Python package is only a ctypes view of those libraries for tests and
bench.py (PyTorch supplies device memory, streams and torch.distributed).  There
is no CPU fallback: importing :mod:`.abi` without the built library raises.
"""
from .abi import (CedError, Context, MultiContext, K7_DEFAULT, K7_TEXTBOOK, Code, lib_path, load_abi,
                  exported_abi_symbols, shard_range)
from .refapi import RefApi

__all__ = ["CedError", "Context", "MultiContext", "shard_range", "Code", "K7_DEFAULT", "K7_TEXTBOOK", "RefApi", "lib_path", "load_abi",
           "exported_abi_symbols"]
