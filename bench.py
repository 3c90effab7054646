#!/usr/bin/env python
"""bench.py -- K=7 r=1/2 Viterbi decoded Gbit/s on B200 (BASELINE.json metric).

One "step" = one pass of the decode hot path (forward ACS + traceback, through the
C ABI ced_decode_batch) over one batch of synthetic frames.  The per-GPU workload
is BASELINE.json configs[1]: 2^16 frames x 4096 information bits, hard decisions,
byte-per-segment symbols, src/defaultParams generators; with N GPUs every rank
decodes its own 2^16 frames (weak scaling, no collective on the data path).

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm
    python bench.py --impl reference [...]                         # reference CPU arm
    python bench.py --mode encode|ber [...]                        # other BASELINE configs

Prints ONE JSON line on rank 0.  `value` is device-timed (CUDA events on the
launching stream) with inputs resident in HBM; `e2e` is the same metric through
ced_decode_batch_host with pinned HOST buffers (H2D + D2H inside the timed
region); `roofline` is the forward ACS kernel against the measured INT-ALU peak;
`cpu_baseline` is the reference's own C decoder (oracle/_ref) on the box's cores;
`per_packet` is the reference driver's own call shape (ONE 2048-bit packet per
synchronous VITERBI_DECODER_HARD call, speedDecode.c:79) through the drop-in host
library, with the reference C on one core beside it in `cpu_baseline`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME_BITS = 4096
FRAMES_PER_GPU = 1 << 16
SEG_STRIDE = 4112                      # 4102 segments padded to a multiple of 16 bytes
INT_OPS_PER_BIT = 256.0 * (FRAME_BITS + 6) / FRAME_BITS        # 256.375
ALGO_BYTES_PER_BIT = ((FRAME_BITS + 6) + FRAME_BITS / 8) / FRAME_BITS   # 1.1265 B / decoded bit
METRIC = "K=7 r=1/2 Viterbi decoded Gbit/s"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p.get("hbm_gbs", 6650.0)), "measured", float(p.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler:
    """SM clock and throttle reasons polled through NVML every few ms DURING the timed region
    (same fields as the nvidia-smi clocks line of B200_PROFILING.md; nvidia-smi itself takes longer
    to start than a timed region lasts)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.power = index, [], set(), []
        self.stop_flag, self.thread, self.max_mhz, self.err = threading.Event(), None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                try:
                    phys = int(visible.split(",")[index])
                except (ValueError, IndexError):
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv, self.err = None, repr(e)

    def _poll(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for name, bit in names.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception as e:  # noqa: BLE001
                self.err = repr(e)
                break
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()

    def stop(self):
        if not self.nv:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: %s" % self.err]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["no samples: %s" % self.err]}
        sm = sorted(self.samples)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_mhz,
                "power_w_max": max(self.power) if self.power else None, "samples": len(sm),
                "reasons": sorted(self.reasons)}


class c_stdout_to_stderr:
    """The reference's init prints a banner with printf (src/viterbiDecoderButterflyk1.c:16);
    keep the process's stdout clean for the single JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        try:
            import ctypes
            ctypes.CDLL(None).fflush(None)
        except Exception:
            pass
        os.dup2(self.saved, 1)
        os.close(self.saved)


def cpu_reference_rate(seconds, threads=None, frames=None):
    with c_stdout_to_stderr():
        return _cpu_reference_rate(seconds, threads, frames)


def _cpu_reference_rate(seconds, threads=None, frames=None):
    """speedDecode-style loop (speedDecode/speedDecode.c:72-110) over the reference's own decoder."""
    import numpy as np
    import oracle
    threads = threads or os.cpu_count() or 1
    frames = frames or max(64, 4 * threads)
    P = oracle.port()
    rng = np.random.default_rng(314)
    msgs = rng.integers(0, 256, (frames, FRAME_BITS // 8), dtype=np.uint8)
    segs = P.encode_batch(7, oracle.K7_G, msgs, seg_stride=SEG_STRIDE)
    flips = rng.random((frames, FRAME_BITS + 6, 2)) < 0.0377
    segs[:, :FRAME_BITS + 6] ^= (flips[..., 0].astype(np.uint8) | (flips[..., 1].astype(np.uint8) << 1))
    R = oracle.ref()
    if R is not None:
        bits, el = R.speed_decode(segs, FRAME_BITS + 6, threads, seconds)
        kind = "reference"
        how = "oracle/_ref (unmodified reference C, -Ofast %s)" % ("x86-64-v4" if R.isa == "v4" else "x86-64-v3")
    else:
        bits, el = P.speed_decode(7, oracle.K7_G, segs, FRAME_BITS + 6, threads, seconds)
        kind = "port"
        how = "oracle/ced_oracle.c (C restatement)"
    return {"value": bits / el / 1e9, "unit": "Gbit/s", "cores": threads, "kind": kind,
            "sample": "%d noisy frames x %d bits decoded round-robin for %.1f s wall on %d threads, %s"
                      % (frames, FRAME_BITS, el, threads, how)}, bits, el


def per_packet_rate(seconds, bits=2048):
    """speedDecode's loop (16 packets, one synchronous last=true call each) through libconvencdec_k7.so."""
    import ctypes
    import numpy as np
    import torch
    import convolutionalencdec_b200 as ced
    with c_stdout_to_stderr():
        api = ced.RefApi("k7")
        enc = api.encoder(); enc.resetConvEncoder(); enc.initConvEncoder()
        dec = api.decoder(); dec.VITERBI_RESET(); dec.VITERBI_INIT()
        rng = np.random.default_rng(314)
        msgs = rng.integers(0, 256, (16, bits // 8), dtype=np.uint8)
        segs = np.stack([enc.convEnc(m, True) for m in msgs])
        u8p = ctypes.POINTER(ctypes.c_uint8)
        out = np.zeros(bits // 8 + 8, dtype=np.uint8)
        ptrs, outp, T = [segs[i].ctypes.data_as(u8p) for i in range(16)], out.ctypes.data_as(u8p), bits + 6
        call = api.lib.viterbiDecoderHardButterflyk1
        lib = ced.load_abi()
        launches0 = int(lib.ced_launch_count(lib.ced_default_ctx()))
        ok = True
        for i in range(64):
            call(dec.p, ptrs[i % 16], outp, T, True)
            ok = ok and bool(np.array_equal(out[:bits // 8], msgs[i % 16]))
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            for i in range(16):
                call(dec.p, ptrs[i], outp, T, True)
            n += 16
        dt = time.perf_counter() - t0
        launches = int(lib.ced_launch_count(lib.ced_default_ctx())) - launches0
    return {"value": n * bits / dt / 1e6, "unit": "Mbit/s", "us_per_call": dt / n * 1e6, "packet_bits": bits,
            "calls": n, "gpu_launches": launches, "round_trip_ok": ok,
            "api": "viterbiDecoderHardButterflyk1(last=true), one packet per synchronous call, host buffers "
                   "(the call speedDecode.c:79 makes); whole packet decoded at once by fpBlockKernel + fpSelectKernel"}


def per_packet_reference_rate(seconds, bits=2048):
    """The same loop on the reference's own C decoder, one host core (oracle/_ref)."""
    import numpy as np
    import oracle
    with c_stdout_to_stderr():
        R = oracle.ref()
        if R is None:
            return None
        rng = np.random.default_rng(314)
        segs = R.encode_batch(rng.integers(0, 256, (16, bits // 8), dtype=np.uint8))
        done, el = R.speed_decode(segs, bits + 6, 1, seconds)
    return {"value": done / el / 1e6, "unit": "Mbit/s", "cores": 1, "kind": "reference"}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    budget = 2.0
    rates, t_all = [], time.time()
    for i in range(args.warmup + args.steps):
        res, bits, el = cpu_reference_rate(budget)
        if i >= args.warmup:
            rates.append((bits, el))
    bits = sum(b for b, _ in rates)
    el = sum(e for _, e in rates)
    value = bits / el / 1e9
    res["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(1, len(rates)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "speedDecode K=7 r=1/2 hard-decision, 4096-bit frames; each step = %.0f s "
                                   "bounded sample on all host cores" % budget,
                       "frame_bits": FRAME_BITS, "channel": "BSC p=0.0377"},
            "cpu_baseline": res,
            "e2e": {"value": value, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.time() - t_all}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="decode", choices=["decode", "encode", "ber"])
    ap.add_argument("--frames", type=int, default=0,
                    help="frames per GPU (default 2^16; 2^20 in encode mode = BASELINE config 3)")
    ap.add_argument("--in-flight", type=int, default=3, help="decode batches in flight (contexts/streams)")
    ap.add_argument("--stride", type=int, default=SEG_STRIDE, help="bytes between frames of the symbol buffer")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.frames <= 0:
        args.frames = (1 << 20) if args.mode == "encode" else FRAMES_PER_GPU

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import convolutionalencdec_b200 as ced
    from convolutionalencdec_b200.sharding import allreduce_counts

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with c_stdout_to_stderr():   # NCCL prints its version banner on stdout at communicator creation
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            warm = torch.zeros(1, device="cuda")
            dist.all_reduce(warm)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = ced.Context(local_rank)
    code = ced.K7_DEFAULT
    frames, bits, T = args.frames, FRAME_BITS, FRAME_BITS + 6
    stream = torch.cuda.Stream()
    first_frame = rank * frames

    # ---- synthetic frames, generated on the device, resident in HBM before timing ----
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    seg_stride = max(args.stride, T)
    segs = torch.zeros((frames, seg_stride), dtype=torch.uint8, device="cuda")
    out = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()   # allocations / zero-fills ran on torch's default stream
    ctx.random_bytes(msgs, seed=314, first_frame=first_frame, stream=stream)
    ctx.encode_batch(code, msgs, out=segs, stream=stream)
    ctx.bsc_channel(segs, T, 2, 0.0377, seed=2718, first_frame=first_frame, stream=stream)
    stream.synchronize()

    hbm_peak, peak_src, _ = load_peaks()
    sampler = ClockSampler(local_rank)

    # Two batches in flight: the forward (ACS) kernel is instruction-issue bound and the traceback kernel is
    # HBM bound, so with one context per stream the traceback of step i overlaps the forward pass of step
    # i+1 (a context serialises its own decodes on its survivor scratch).  `value` is this steady-state
    # throughput; the one-decode-at-a-time figure is reported as `single_stream`.
    extra = [ced.Context(local_rank) for _ in range(max(1, args.in_flight) - 1)]
    lanes = [(ctx, stream, out)] + [(c_, torch.cuda.Stream(), torch.empty_like(out)) for c_ in extra]

    if args.mode == "encode":
        def step(i=0):
            ctx.encode_batch(code, msgs, out=segs, stream=stream)
        units = frames * bits
        lanes = lanes[:1]
    elif args.mode == "ber":
        counters = torch.zeros(4, dtype=torch.int64, device="cuda")

        def step(i=0):
            ctx.encode_batch(code, msgs, out=segs, stream=stream)
            ctx.bsc_channel(segs, T, 2, 0.0377, seed=2718, first_frame=first_frame, counters=counters[:2],
                            stream=stream)
            ctx.decode_batch(code, segs, bits, out=out, stream=stream)
            ctx.ber_count(out, msgs, counters[2:], stream=stream)
        units = frames * bits
        lanes = lanes[:1]
    else:
        def step(i=0):
            c_, s_, o_ = lanes[i % len(lanes)]
            c_.decode_batch(code, segs, bits, out=o_, stream=s_)
        units = frames * bits

    def timed_run(n_steps, n_lanes):
        """n_steps steps round-robin over n_lanes (context, stream) pairs; device time start -> all done."""
        use = lanes[:n_lanes]
        timing = torch.cuda.Stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(timing)
        for _, s_, _ in use:
            s_.wait_event(e0)
        for i in range(n_steps):
            c_, s_, o_ = use[i % n_lanes]
            if args.mode == "decode":
                c_.decode_batch(code, segs, bits, out=o_, stream=s_)
            else:
                step(i)
        for _, s_, _ in use:
            done = torch.cuda.Event()
            done.record(s_)
            timing.wait_event(done)
        e1.record(timing)
        timing.synchronize()
        return e0.elapsed_time(e1)

    for i in range(args.warmup * len(lanes)):
        step(i)
    torch.cuda.synchronize()
    barrier()
    launches0 = sum(c_.launches for c_, _, _ in lanes)
    sampler.start()
    t_wall = time.perf_counter()
    ms_total_local = timed_run(args.steps, len(lanes))
    barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    launches = sum(c_.launches for c_, _, _ in lanes) - launches0
    ms_total = torch.tensor([ms_total_local], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_total = float(ms_total.item())
    ms_per_step = ms_total / args.steps
    value = world * units / (ms_per_step * 1e-3) / 1e9
    single = None
    if len(lanes) > 1:
        barrier()
        ms1 = torch.tensor([timed_run(args.steps, 1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms1, op=dist.ReduceOp.MAX)
        single = {"value": world * units * args.steps / (float(ms1.item()) * 1e-3) / 1e9, "unit": "Gbit/s",
                  "ms_per_step": float(ms1.item()) / args.steps,
                  "note": "one decode at a time on one stream (forward then traceback, no overlap)"}

    line = {"metric": METRIC if args.mode == "decode" else
            ("K=7 r=1/2 convolutional encoded Gbit/s (information bits)" if args.mode == "encode" else
             "K=7 r=1/2 BER pipeline Gbit/s (encode+BSC+decode+count)"),
            "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": ("speedEncode K=7 rate-1/2 (0113/0171), %d frames x 4096 bits per GPU" % frames)
                       if args.mode == "encode" else
                       ("speedDecode K=7 rate-1/2 (0113/0171) hard-decision, 2^16 frames x 4096 bits per GPU"
                        if frames == FRAMES_PER_GPU else
                        "K=7 rate-1/2 hard-decision, %d frames x 4096 bits per GPU" % frames),
                       "mode": args.mode, "frames_per_gpu": frames, "frame_bits": bits, "segment_stride_bytes": seg_stride,
                       "symbol_format": "1 byte per 2-bit segment (reference wire format)", "channel": "BSC p=0.0377 (Eb/N0 5 dB)",
                       "l2_policy": "inputs (%.0f MB symbols + %.0f MB survivors per step) exceed the 126 MB L2"
                                    % (frames * seg_stride / 1e6, frames * (T // 2) * 16 / 1e6),
                       "sharding": "frames [rank*F, (rank+1)*F) per rank, no data-path collective"},
            "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall}
    if single is not None:
        line["single_stream"] = single
        line["config"]["in_flight"] = "%d batches (one ced_ctx + CUDA stream each): traceback(i) overlaps forward(i+1)" % len(lanes)

    if args.mode == "decode":
        # ---- roofline of the dominant kernel (forward ACS), CUDA events around that kernel alone ----
        ctx.set_profiling(True)
        fwd, tb = [], []
        for _ in range(max(3, min(args.steps, 10))):
            ctx.decode_batch(code, segs, bits, out=out, stream=stream)
            f, t = ctx.last_kernel_ms()
            fwd.append(f)
            tb.append(t)
        ctx.set_profiling(False)
        fwd_ms, tb_ms = sum(fwd) / len(fwd), sum(tb) / len(tb)
        int_peak = ctx.probe_int_peak(0)
        int_peak_dual = ctx.probe_int_peak(1)
        algo_ops = units * INT_OPS_PER_BIT
        achieved = algo_ops / (fwd_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("k7ForwardKernel_dram_bytes_per_launch")
        # instruction-level view: the 6-step loop body of k7ForwardKernel is 724 SASS instructions (cuobjdump -sass
        # of the committed build) = 120.7 warp-instructions per 32 frame-steps; an SM sub-partition
        # issues at most one warp-instruction per cycle and the kernel splits them ~50/50 over the ALU and FMA pipes
        sm_mhz = (clocks.get("sm_mhz") or 1965.0)
        warp_instr = frames * T * (724.0 / 6.0) / 32.0
        ipc = warp_instr / (148 * 4 * fwd_ms * 1e-3 * sm_mhz * 1e6)
        line["roofline"] = {"bound": "int_alu", "kernel": "k7ForwardKernel", "achieved": achieved,
                            "issue": {"instr_per_frame_step": 724.0 / 6.0, "ipc_per_sm_subpartition": ipc, "peak": 1.0,
                                      "frac": ipc, "note": "SASS instruction count x frame-steps / (592 sub-partitions x "
                                                            "kernel cycles at the sampled SM clock)"},
                            "peak": int_peak / 1e12, "unit": "Tiop/s", "frac": achieved / (int_peak / 1e12),
                            "peak_source": "measured live: dependent-free LOP3 stream (ced_probe_int_peak mode 0)",
                            "peak_with_imad_coissue": int_peak_dual / 1e12,
                            "algorithmic_ops_per_launch": algo_ops, "kernel_ms": fwd_ms,
                            "kernel_share_of_step": fwd_ms / (fwd_ms + tb_ms), "traceback_ms": tb_ms,
                            "traffic": traffic,
                            "hbm": {"achieved": units * ALGO_BYTES_PER_BIT / ((fwd_ms + tb_ms) * 1e-3) / 1e9,
                                    "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                                    "frac": units * ALGO_BYTES_PER_BIT / ((fwd_ms + tb_ms) * 1e-3) / 1e9 / hbm_peak,
                                    "note": "algorithmic symbol-in + bits-out bytes; not the binding roofline"}}

    if args.mode == "encode":
        # the encoder is HBM-bound: message bytes in + one byte per coded segment out
        enc_bytes = frames * (bits // 8 + T)
        achieved = enc_bytes / (ms_per_step * 1e-3) / 1e9
        enc_traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath) and frames == 1 << 20:
            with open(tpath) as f:
                enc_traffic = json.load(f).get("encodeBatchLutKernel_dram_bytes_per_launch_2p20_frames")
        line["roofline"] = {"bound": "hbm", "kernel": "encodeBatchLutKernel", "achieved": achieved, "peak": hbm_peak,
                            "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": enc_bytes, "kernel_ms": ms_per_step, "traffic": enc_traffic}

    def host_calls_in_flight(call, n_calls, n_threads):
        """n_calls synchronous host-buffer calls issued from n_threads host threads (one ced_ctx and one set of
        pinned buffers each, ctypes releases the GIL): while one call drains its pipeline the other one's H2D
        copies keep the PCIe link busy.  Returns wall seconds, max over ranks."""
        import threading
        per = [n_calls // n_threads + (1 if i < n_calls % n_threads else 0) for i in range(n_threads)]

        def worker(i):
            torch.cuda.set_device(local_rank)
            for _ in range(per[i]):
                call(i)
        barrier()
        t0 = time.perf_counter()
        threads = [threading.Thread(target=worker, args=(i,)) for i in range(n_threads)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        el = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
        return float(el.item())

    if not args.no_e2e and args.mode == "decode":
        # ---- e2e: the public host-buffer call; H2D of the symbols and D2H of the bits inside the timed region ----
        n_host = 2
        host_ctx = [ctx] + [ced.Context(local_rank) for _ in range(n_host - 1)]
        h_segs = [torch.empty((frames, seg_stride), dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        h_out = [torch.empty((frames, bits // 8), dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        for h in h_segs:
            h.copy_(segs)
        torch.cuda.synchronize()
        n_e2e = max(4, min(args.steps, 8))
        # warm-up: the library measures during its first 16 host-buffer calls per device whether packing part of
        # the chunks on the host beats raw copies on this machine; the timed calls come after that
        host_calls_in_flight(lambda i: host_ctx[i].decode_batch_host(code, h_segs[i], bits, h_out[i]), 20, n_host)
        l0 = sum(c_.launches for c_ in host_ctx)
        el = host_calls_in_flight(lambda i: host_ctx[i].decode_batch_host(code, h_segs[i], bits, h_out[i]), n_e2e, n_host)
        el1 = host_calls_in_flight(lambda i: host_ctx[0].decode_batch_host(code, h_segs[0], bits, h_out[0]), n_e2e, 1)
        ok = all(bool(torch.equal(h.cuda(), out)) for h in h_out)
        line["e2e"] = {"value": world * units * n_e2e / el / 1e9, "unit": "Gbit/s",
                       "h2d_bytes_per_step": (frames - 1) * seg_stride + T,
                       "d2h_bytes_per_step": frames * bits // 8, "steps": n_e2e,
                       "api": "ced_decode_batch_host (pinned host buffers, 8192-frame chunks, 4 in flight: H2D, 4 compute "
                              "streams, D2H; host worker threads pack to 2 bits the chunks the copy engine is not ready "
                              "for, where 16 calibration calls showed that to be faster than raw copies on this "
                              "machine); %d host threads, one context each, keep calls in flight" % n_host,
                       "one_call_at_a_time": world * units * n_e2e / el1 / 1e9,
                       "matches_device_path": ok, "gpu_launches": sum(c_.launches for c_ in host_ctx) - l0}

    if not args.no_e2e and args.mode == "decode":
        # ---- same call on the packed wire format (4 segments per byte; not a reference format) ----
        pstride = ((T + 3) // 4 + 15) // 16 * 16
        d_packed = ctx.pack_symbols(segs, T, packed_stride=pstride, stream=stream)
        stream.synchronize()
        h_packed = torch.empty((frames, pstride), dtype=torch.uint8).pin_memory()
        h_packed.copy_(d_packed)
        torch.cuda.synchronize()
        h_packed2 = [h_packed] + [h_packed.clone().pin_memory() for _ in range(n_host - 1)]
        for i in range(n_host):
            host_ctx[i].decode_batch_packed_host(code, h_packed2[i], bits, h_out[i])
        el = host_calls_in_flight(lambda i: host_ctx[i].decode_batch_packed_host(code, h_packed2[i], bits, h_out[i]),
                                  n_e2e, n_host)
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out_p = torch.empty_like(out)
        ctx.decode_batch_packed(code, d_packed, bits, out=out_p, stream=stream)
        ev2.record(stream)
        for _ in range(n_e2e):
            ctx.decode_batch_packed(code, d_packed, bits, out=out_p, stream=stream)
        ev3.record(stream)
        stream.synchronize()
        line["packed_format"] = {"e2e": {"value": world * units * n_e2e / el / 1e9, "unit": "Gbit/s",
                                         "h2d_bytes_per_step": (frames - 1) * pstride + (T + 3) // 4,
                                         "d2h_bytes_per_step": frames * bits // 8,
                                         "api": "ced_decode_batch_packed_host"},
                                 "device_resident_value": world * units * n_e2e / (ev2.elapsed_time(ev3) * 1e-3) / 1e9,
                                 "matches_byte_format": bool(all(torch.equal(h.cuda(), out) for h in h_out) and torch.equal(out_p, out)),
                                 "note": "4 two-bit segments per byte; not the reference wire format, reported beside it"}

    # ---- decoded bit-error count, summed over ranks with NCCL (BER mode's only collective) ----
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    if args.mode == "encode":   # nothing was decoded in this mode: decode the freshly encoded (noise-free) symbols
        ctx.decode_batch(code, segs, bits, out=out, stream=stream)
    ctx.ber_count(out, msgs, cnt, stream=stream)
    stream.synchronize()
    allreduce_counts(cnt)
    line["check"] = {"decoded_bit_errors": int(cnt[0].item()), "decoded_bits": int(cnt[1].item()),
                     "ber": float(cnt[0].item()) / max(1, int(cnt[1].item()))}

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.mode == "decode":
        line["cpu_baseline"] = cpu_reference_rate(4.0)[0]
        # the reference's own driver shape: ONE 2048-bit packet per synchronous VITERBI_DECODER_HARD call
        # (speedDecode.c:18-23,79) through the drop-in host C library, next to the reference C on one core
        line["per_packet"] = per_packet_rate(1.0)
        line["cpu_baseline"]["one_core_2048_bit_packets"] = per_packet_reference_rate(1.0)

    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    for c_ in extra:
        c_.close()
    if not args.no_e2e and args.mode == "decode":
        for c_ in host_ctx[1:]:
            c_.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
