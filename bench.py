#!/usr/bin/env python
"""bench.py -- K=7 r=1/2 Viterbi decoded Gbit/s on B200 (BASELINE.json metric).

One "step" = one pass of the decode hot path (forward ACS + traceback, through the C ABI ced_decode_batch)
over this rank's synthetic frames.
  * --gpus 1 (default): BASELINE.json configs[1] -- 2^16 frames x 4096 information bits, hard decisions,
    byte-per-segment symbols, src/defaultParams generators.  The same line carries sub-records for the other
    configs: `encode` (configs[2], 2^20 frames), `ber` (configs[3] pipeline + a BER point) and `soft`
    (soft-decision decoder).
  * --gpus N > 1 (under torchrun, one rank per GPU): BASELINE.json configs[4] -- 2^22 frames sharded,
    2^22 / N per rank, "scaling": "strong"; no collective on the data path; one ced_decode_batch call per
    step over the rank's shard, two steps in flight.  The weak figure (2^16 frames per GPU) is kept in `weak`.

    python bench.py [--gpus N] [--steps K] [--warmup W]           # our arm
    python bench.py --impl reference [...]                         # reference CPU arm
    python bench.py --mode encode|ber [...]                        # other BASELINE configs as the main line

Prints ONE JSON line on rank 0.  `value` is device-timed (CUDA events on the launching stream) with inputs
resident in HBM; `e2e` is the same metric through ced_decode_batch_host with pinned HOST buffers (H2D + D2H inside
the timed region), with the raw pinned-copy ceiling of the same ranks beside it; `roofline` is the forward ACS
kernel against the measured INT-ALU peak; `cpu_baseline` is the reference's own C decoder (oracle/_ref) on the
box's cores; `per_packet` is the reference driver's own call shape (ONE 2048-bit packet per synchronous
VITERBI_DECODER_HARD call, speedDecode.c:79) through the drop-in host library.
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
FRAMES_PER_GPU = 1 << 16               # BASELINE configs[1]; also the sub-batch a rank's shard is decoded in
CONFIG5_FRAMES = 1 << 22               # BASELINE configs[4]: sharded over the ranks (strong scaling)
CONFIG3_FRAMES = 1 << 20               # BASELINE configs[2]: encoder
SEG_STRIDE = 4112                      # 4102 segments padded to a multiple of 16 bytes
INT_OPS_PER_BIT = 256.0 * (FRAME_BITS + 6) / FRAME_BITS        # 256.375
ALGO_BYTES_PER_BIT = ((FRAME_BITS + 6) + FRAME_BITS / 8) / FRAME_BITS   # 1.1265 B / decoded bit
METRIC = "K=7 r=1/2 Viterbi decoded Gbit/s"


def load_sass_stats():
    """Instruction counts of the hot loops of the library that is loaded, written by `make cuda`
    (tools/sass_loop_stats.py disassembles the built libced_cuda.so)."""
    path = os.path.join(ROOT, "convolutionalencdec_b200", "sass_stats.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


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
            "sample": "%d noisy frames x %d bits decoded round-robin for %.1f s wall on %d threads (one pinned per CPU), %s"
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
                   "(the call speedDecode.c:79 makes); the packet's 64-step blocks decoded at once by wsBlockKernel (speculative start "
                   "metrics), checked, repaired and walked back by wsJoinKernel (csrc/warp_split.cu)"}


def small_batch_rate(ctx, seconds, packets=16, bits=2048):
    """speedDecode's shape -- 16 packets of 2048 bits (speedDecode.c:18-19) -- as ONE ced_decode_batch_host call from
    pageable host buffers: one copy in, the small-batch kernels (csrc/warp_split.cu, warp_frame.cu), one copy out."""
    import numpy as np
    import convolutionalencdec_b200 as ced
    rng = np.random.default_rng(314)
    msgs = rng.integers(0, 256, (packets, bits // 8), dtype=np.uint8)
    segs = np.zeros((packets, bits + 6), dtype=np.uint8)
    ctx.encode_batch_host(ced.K7_DEFAULT, msgs, segs)
    out = np.zeros((packets, bits // 8), dtype=np.uint8)
    launches0 = ctx.launches
    for _ in range(10):
        ctx.decode_batch_host(ced.K7_DEFAULT, segs, bits, out)
    ok = bool(np.array_equal(out, msgs))
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        ctx.decode_batch_host(ced.K7_DEFAULT, segs, bits, out)
        n += 1
    dt = time.perf_counter() - t0
    return {"value": n * packets * bits / dt / 1e6, "unit": "Mbit/s", "us_per_call": dt / n * 1e6, "packets_per_call": packets,
            "packet_bits": bits, "calls": n, "gpu_launches": ctx.launches - launches0, "round_trip_ok": ok,
            "api": "ced_decode_batch_host, pageable host buffers, one synchronous call per %d packets: packets cut into "
                   "blocks in time, one warp per block from speculative start metrics (radix-4 steps), hand-overs checked and "
                   "wrong guesses re-run by the join kernel, warp-parallel traceback" % packets}


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


def bind_to_gpu_numa_node(local_rank):
    """One process per GPU: run this rank's host threads -- and therefore place its page-locked buffers, which are
    allocated first-touch -- on the CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI device).
    Without it every rank's host buffers may sit on one socket and the other socket's GPUs pull their symbols across
    the inter-socket link (round 1: e2e flat from 2 to 4 GPUs).  CED_BENCH_NUMA=0 turns it off.  Returns a note."""
    if os.environ.get("CED_BENCH_NUMA", "1") == "0":
        return "off (CED_BENCH_NUMA=0)"
    try:
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        bus = nv.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:      # nvml prints an 8-digit domain, sysfs a 4-digit one
            bus = bus[4:]
        base = "/sys/bus/pci/devices/" + bus
        with open(base + "/local_cpulist") as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        node = "?"
        try:
            with open(base + "/numa_node") as f:
                node = f.read().strip()
        except OSError:
            pass
        if not use or len(use) == len(allowed):
            return "node %s: all %d allowed CPUs are local" % (node, len(allowed))
        os.sched_setaffinity(0, use)
        return "node %s: bound to %d of %d CPUs (%s)" % (node, len(use), len(allowed), text)
    except Exception as e:  # noqa: BLE001
        return "unavailable (%s)" % type(e).__name__


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="decode", choices=["decode", "encode", "ber"])
    ap.add_argument("--workload", default="auto", choices=["auto", "config2", "config5"],
                    help="auto: BASELINE configs[1] (2^16 frames) on one GPU, configs[4] (2^22 frames sharded, strong "
                         "scaling) on several")
    ap.add_argument("--frames", type=int, default=0,
                    help="frames per GPU (overrides --workload; 2^20 in encode mode = BASELINE config 3)")
    ap.add_argument("--in-flight", type=int, default=3, help="decode batches in flight (contexts/streams)")
    ap.add_argument("--stride", type=int, default=SEG_STRIDE, help="bytes between frames of the symbol buffer")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-subrecords", action="store_true", help="skip the encode / ber / soft sub-records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return

    workload = args.workload
    if workload == "auto":
        workload = "config5" if (world > 1 and args.mode == "decode" and args.frames <= 0) else "config2"
    strong = workload == "config5" and args.frames <= 0
    if args.frames <= 0:
        args.frames = CONFIG3_FRAMES if args.mode == "encode" else (CONFIG5_FRAMES // world if strong else FRAMES_PER_GPU)

    import numpy as np
    import torch
    import torch.distributed as dist
    import convolutionalencdec_b200 as ced
    from convolutionalencdec_b200.sharding import allreduce_counts

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (there is no CPU fallback); use --impl reference for the CPU arm")
    numa_note = bind_to_gpu_numa_node(local_rank)   # before the CUDA context and any page-locked allocation
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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    ctx = ced.Context(local_rank)
    code = ced.K7_DEFAULT
    frames, bits, T = args.frames, FRAME_BITS, FRAME_BITS + 6
    stream = torch.cuda.Stream()
    first_frame = rank * frames
    sass = load_sass_stats()

    # ---- synthetic frames, generated on the device, resident in HBM before timing ----
    # config 2: one ced_decode_batch call of 2^16 frames per step.  config 5 (strong scaling): ONE call over the rank's
    # whole shard per step -- the library keeps waves of 2^16 frames in flight internally (decodeBatchPipelined).
    # `sub` = frames per call of a step, `launch` = frames of the single-wave calls used for the roofline / e2e / weak legs
    launch = min(frames, FRAMES_PER_GPU)
    sub = frames if strong else launch
    n_sub = (frames + sub - 1) // sub
    msgs = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    seg_stride = max(args.stride, T)
    segs = torch.zeros((frames, seg_stride), dtype=torch.uint8, device="cuda")
    out = torch.empty((frames, bits // 8), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()   # allocations / zero-fills ran on torch's default stream
    for a in range(0, frames, launch):
        b = min(frames, a + launch)
        ctx.random_bytes(msgs[a:b], seed=314, first_frame=first_frame + a, stream=stream)
        ctx.encode_batch(code, msgs[a:b], out=segs[a:b], stream=stream)
        ctx.bsc_channel(segs[a:b], T, 2, 0.0377, seed=2718, first_frame=first_frame + a, stream=stream)
    stream.synchronize()

    hbm_peak, peak_src, _ = load_peaks()
    sampler = ClockSampler(local_rank)

    # Batches in flight: the forward (ACS) kernel is instruction-issue bound and the traceback kernel is HBM bound,
    # so with one context per stream the traceback of sub-batch i overlaps the forward pass of sub-batch i+1 (a
    # context serialises its own decodes on its survivor scratch).  `value` is this steady-state throughput; the
    # one-decode-at-a-time figure is reported as `single_stream`.
    extra = [ced.Context(local_rank) for _ in range(max(1, args.in_flight) - 1)]
    all_lanes = [(ctx, stream)] + [(c_, torch.cuda.Stream()) for c_ in extra]
    # config 5: one call covers the shard and keeps its waves in flight by itself; two calls (steps) in flight so that
    # the ramp of a call -- first wave's forward pass alone, last wave's traceback alone -- overlaps the neighbouring step
    lanes = all_lanes[:2] if strong else all_lanes
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")

    # one sub-batch per step: every lane decodes into its own output buffer (the same frames are in flight on several
    # lanes at once); several sub-batches per step: each writes its own rows of `out`
    lane_out = {id(ctx): out}
    if n_sub == 1:
        for c_, _s in lanes[1:]:
            lane_out[id(c_)] = torch.empty_like(out)

    def decode_pass(use, first_lane=0):
        """one step: every sub-batch of this rank's frames, round-robin over the (context, stream) lanes"""
        for j in range(n_sub):
            c_, s_ = use[(first_lane + j) % len(use)]
            a, b = j * sub, min(frames, (j + 1) * sub)
            c_.decode_batch(code, segs[a:b], bits, out=lane_out.get(id(c_), out)[a:b], stream=s_)

    def other_step():
        if args.mode == "encode":
            ctx.encode_batch(code, msgs, out=segs, stream=stream)
        else:
            ctx.encode_batch(code, msgs, out=segs, stream=stream)
            ctx.bsc_channel(segs, T, 2, 0.0377, seed=2718, first_frame=first_frame, counters=counters[:2], stream=stream)
            ctx.decode_batch(code, segs, bits, out=out, stream=stream)
            ctx.ber_count(out, msgs, counters[2:], stream=stream)

    if args.mode != "decode":
        lanes = lanes[:1]
    units = frames * bits

    def timed_run(n_steps, n_lanes):
        """n_steps steps over n_lanes (context, stream) pairs; device time from start to all done."""
        use = lanes[:n_lanes]
        timing = torch.cuda.Stream()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(timing)
        for _, s_ in use:
            s_.wait_event(e0)
        for i in range(n_steps):
            if args.mode == "decode":
                decode_pass(use, first_lane=i * n_sub)
            else:
                other_step()
        for _, s_ in use:
            done = torch.cuda.Event()
            done.record(s_)
            timing.wait_event(done)
        e1.record(timing)
        timing.synchronize()
        return e0.elapsed_time(e1)

    for i in range(args.warmup):
        if args.mode == "decode":
            decode_pass(lanes, first_lane=i * n_sub)
            if n_sub == 1:            # every context warms up (scratch allocation) even with one sub-batch per step
                for l in range(1, len(lanes)):
                    decode_pass(lanes[l:l + 1])
        else:
            other_step()
    torch.cuda.synchronize()
    barrier()
    launches0 = sum(c_.launches for c_, _ in lanes)
    sampler.start()
    t_wall = time.perf_counter()
    ms_total_local = timed_run(args.steps, len(lanes))
    barrier()
    wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    launches = sum(c_.launches for c_, _ in lanes) - launches0
    ms_total = max_over_ranks(ms_total_local)
    ms_per_step = ms_total / args.steps
    value = world * units / (ms_per_step * 1e-3) / 1e9
    single = None
    if len(lanes) > 1:
        barrier()
        ms1 = max_over_ranks(timed_run(args.steps, 1))
        single = {"value": world * units * args.steps / (ms1 * 1e-3) / 1e9, "unit": "Gbit/s",
                  "ms_per_step": ms1 / args.steps,
                  "note": ("one ced_decode_batch call at a time over the whole shard (waves pipelined inside the call)" if strong
                           else "one ced_decode_batch at a time on one context and stream (forward then traceback, no overlap)")}

    total_frames = world * frames
    if args.mode == "encode":
        wl = "speedEncode K=7 rate-1/2 (0113/0171), %d frames x 4096 bits per GPU" % frames
    elif strong:
        wl = ("speedDecode K=7 rate-1/2 (0113/0171) hard-decision, 2^22 frames x 4096 bits sharded over %d GPU%s "
              "(BASELINE configs[4])" % (world, "" if world == 1 else "s"))
    elif frames == FRAMES_PER_GPU:
        wl = "speedDecode K=7 rate-1/2 (0113/0171) hard-decision, 2^16 frames x 4096 bits per GPU (BASELINE configs[1])"
    else:
        wl = "K=7 rate-1/2 hard-decision, %d frames x 4096 bits per GPU" % frames
    line = {"metric": METRIC if args.mode == "decode" else
            ("K=7 r=1/2 convolutional encoded Gbit/s (information bits)" if args.mode == "encode" else
             "K=7 r=1/2 BER pipeline Gbit/s (encode+BSC+decode+count)"),
            "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl, "mode": args.mode, "frames_per_gpu": frames, "total_frames": total_frames,
                       "frame_bits": bits, "segment_stride_bytes": seg_stride,
                       "calls_per_step": n_sub if args.mode == "decode" else 1,
                       "symbol_format": "1 byte per 2-bit segment (reference wire format)", "channel": "BSC p=0.0377 (Eb/N0 5 dB)",
                       "l2_policy": "inputs (%.0f MB symbols + %.0f MB survivors per wave of 2^16 frames) exceed the 126 MB L2"
                                    % (launch * seg_stride / 1e6, launch * (T // 2) * 16 / 1e6),
                       "sharding": "frames [rank*F, (rank+1)*F) per rank, no data-path collective",
                       "host_numa": numa_note},
            "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall}
    if single is not None:
        line["single_stream"] = single
    if single is not None and not strong:
        line["config"]["in_flight"] = ("%d ced_decode_batch calls in flight (one ced_ctx + CUDA stream each): "
                                       "traceback(i) overlaps forward(i+1)" % len(lanes))
    elif strong:
        line["config"]["in_flight"] = ("one ced_decode_batch call per step over the whole shard; the library keeps 3 waves of "
                                       "2^16 frames in flight on internal streams (decodeBatchPipelined); %d steps in flight "
                                       "(one ced_ctx + stream each)" % len(lanes))

    if args.mode == "decode":
        # ---- roofline of the dominant kernel (forward ACS), CUDA events around that kernel alone ----
        ctx.set_profiling(True)
        fwd, tb = [], []
        for _ in range(max(3, min(args.steps, 10))):
            ctx.decode_batch(code, segs[:launch], bits, out=out[:launch], stream=stream)
            f, t = ctx.last_kernel_ms()
            fwd.append(f)
            tb.append(t)
        ctx.set_profiling(False)
        fwd_ms, tb_ms = sum(fwd) / len(fwd), sum(tb) / len(tb)
        int_peak = ctx.probe_int_peak(0)
        int_peak_dual = ctx.probe_int_peak(1)
        launch_bits = launch * bits              # one launch = 2^16 frames
        algo_ops = launch_bits * INT_OPS_PER_BIT
        achieved = algo_ops / (fwd_ms * 1e-3) / 1e12
        # DRAM bytes of one launch: ncu --set full capture at 2^16 frames (profiles/roofline_traffic.json), scaled
        # to this launch's frame count -- the traffic is survivor stores + symbol loads, both linear in frames
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            per_launch = tj.get("k7ForwardKernel_dram_bytes_per_launch")
            if per_launch:
                traffic = per_launch * launch / float(tj.get("frames_per_launch", FRAMES_PER_GPU))
        # instruction-level view: SASS instruction count of the kernel's 6-step loop body, read from the BUILT
        # library by tools/sass_loop_stats.py at `make cuda`; an SM sub-partition issues at most one warp-instruction
        # per cycle and the kernel splits them ~50/50 over the ALU and FMA pipes
        loop = sass.get("k7_forward", {})
        ipf = loop.get("instr_per_frame_step")
        sm_mhz = (clocks.get("sm_mhz") or 1965.0)
        issue = None
        if ipf:
            warp_instr = launch * T * ipf / 32.0
            ipc = warp_instr / (148 * 4 * fwd_ms * 1e-3 * sm_mhz * 1e6)
            issue = {"instr_per_frame_step": ipf, "loop_instructions": loop.get("loop_instructions"),
                     "pipes": loop.get("pipes"), "ipc_per_sm_subpartition": ipc, "peak": 1.0, "frac": ipc,
                     "source": "cuobjdump -sass of the built libced_cuda.so (convolutionalencdec_b200/sass_stats.json, "
                               "profiles/k7_forward_loop.sass)",
                     "note": "SASS instruction count x frame-steps / (592 sub-partitions x kernel cycles at the "
                             "sampled SM clock)"}
        line["roofline"] = {"bound": "int_alu", "kernel": "k7ForwardKernel", "achieved": achieved, "issue": issue,
                            "peak": int_peak / 1e12, "unit": "Tiop/s", "frac": achieved / (int_peak / 1e12),
                            "peak_source": "measured live: dependent-free LOP3 stream (ced_probe_int_peak mode 0)",
                            "peak_with_imad_coissue": int_peak_dual / 1e12,
                            "algorithmic_ops_per_launch": algo_ops, "frames_per_launch": launch, "kernel_ms": fwd_ms,
                            "kernel_share_of_step": fwd_ms / (fwd_ms + tb_ms), "traceback_ms": tb_ms,
                            "traffic": traffic,
                            "hbm": {"achieved": launch_bits * ALGO_BYTES_PER_BIT / ((fwd_ms + tb_ms) * 1e-3) / 1e9,
                                    "peak": hbm_peak, "unit": "GB/s", "peak_source": peak_src,
                                    "frac": launch_bits * ALGO_BYTES_PER_BIT / ((fwd_ms + tb_ms) * 1e-3) / 1e9 / hbm_peak,
                                    "note": "algorithmic symbol-in + bits-out bytes; not the binding roofline"}}

    if args.mode == "encode":
        line["roofline"] = encode_roofline(frames, bits, T, ms_per_step, hbm_peak, peak_src)

    if strong and args.mode == "decode":
        # ---- the weak figure beside the strong one: 2^16 frames per GPU, 3 calls in flight (BASELINE configs[1] per GPU) ----
        lanes_w = all_lanes
        weak_out = [torch.empty((launch, bits // 8), dtype=torch.uint8, device="cuda") for _ in lanes_w]

        def weak_pass(n_steps):
            timing = torch.cuda.Stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(timing)
            for _, s_ in lanes_w:
                s_.wait_event(e0)
            for i in range(n_steps):
                c_, s_ = lanes_w[i % len(lanes_w)]
                c_.decode_batch(code, segs[:launch], bits, out=weak_out[i % len(lanes_w)], stream=s_)
            for _, s_ in lanes_w:
                done = torch.cuda.Event()
                done.record(s_)
                timing.wait_event(done)
            e1.record(timing)
            timing.synchronize()
            return e0.elapsed_time(e1)
        weak_pass(2 * len(lanes_w))
        barrier()
        n_w = max(10, min(args.steps, 30))
        ms_w = max_over_ranks(weak_pass(n_w))
        line["weak"] = {"value": world * launch * bits * n_w / (ms_w * 1e-3) / 1e9, "unit": "Gbit/s", "frames_per_gpu": launch,
                        "steps": n_w, "note": "2^16 frames per GPU, %d calls in flight: the N = 1 workload on every GPU" % len(lanes_w)}

    def host_calls_in_flight(call, n_calls, n_threads):
        """n_calls synchronous host-buffer calls issued from n_threads host threads (one ced_ctx and one set of
        pinned buffers each, ctypes releases the GIL): while one call drains its pipeline the other one's H2D
        copies keep the PCIe link busy.  Returns wall seconds, max over ranks."""
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
        return max_over_ranks(time.perf_counter() - t0)

    host_ctx = []
    if not args.no_e2e and args.mode == "decode":
        # ---- e2e: the public host-buffer call; H2D of the symbols and D2H of the bits inside the timed region.
        # Host batch = 2^16 frames per rank at every N (BASELINE configs[1] per GPU): page-locking 2^22 frames of
        # symbols (17 GB) per box would measure the host's memory, not the library.
        n_host = 2
        hf = launch
        host_ctx = [ctx] + [ced.Context(local_rank) for _ in range(n_host - 1)]
        h_segs = [torch.empty((hf, seg_stride), dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        h_out = [torch.empty((hf, bits // 8), dtype=torch.uint8).pin_memory() for _ in range(n_host)]
        for h in h_segs:
            h.copy_(segs[:hf])
        torch.cuda.synchronize()
        n_e2e = max(4, min(args.steps, 8))
        # raw pinned-copy ceiling of the same ranks, all at once: what the links + host memory give with no kernels
        barrier()
        up, down = ctx.probe_copy_ceiling(256 << 20, 3)
        h2d_ceiling = sum_over_ranks(up) / 1e9
        d2h_ceiling = sum_over_ranks(down) / 1e9
        # warm-up: the library measures during its first 16 host-buffer calls per device whether packing part of
        # the chunks on the host beats raw copies on this machine; the timed calls come after that
        host_calls_in_flight(lambda i: host_ctx[i].decode_batch_host(code, h_segs[i], bits, h_out[i]), 20, n_host)
        l0 = sum(c_.launches for c_ in host_ctx)
        el = host_calls_in_flight(lambda i: host_ctx[i].decode_batch_host(code, h_segs[i], bits, h_out[i]), n_e2e, n_host)
        el1 = host_calls_in_flight(lambda i: host_ctx[0].decode_batch_host(code, h_segs[0], bits, h_out[0]), n_e2e, 1)
        ok = all(bool(torch.equal(h.cuda(), out[:hf])) for h in h_out)
        h2d_bytes = (hf - 1) * seg_stride + T
        e2e_value = world * hf * bits * n_e2e / el / 1e9
        sym_gbs = world * h2d_bytes * n_e2e / el / 1e9       # rate at which the caller's symbol bytes were consumed
        line["e2e"] = {"value": e2e_value, "unit": "Gbit/s",
                       "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": hf * bits // 8, "steps": n_e2e,
                       "frames_per_call": hf,
                       "h2d_ceiling_gbs": h2d_ceiling, "d2h_ceiling_gbs": d2h_ceiling,
                       "symbol_gbs": sym_gbs, "frac_of_ceiling": sym_gbs / h2d_ceiling if h2d_ceiling else None,
                       "ceiling_note": "raw cudaMemcpyAsync of 256 MiB page-locked buffers on all %d rank(s) at once, best of 3, "
                                       "summed (ced_probe_copy_ceiling); symbol_gbs = caller symbol bytes consumed per second -- "
                                       "above 1.0 of the ceiling only where host threads pack part of the chunks to 2 bits "
                                       "before the copy" % world,
                       "api": "ced_decode_batch_host (pinned host buffers, 8192-frame chunks, 4 in flight: H2D, 4 compute "
                              "streams, D2H; host worker threads pack to 2 bits the chunks the copy engine is not ready "
                              "for, where 16 calibration calls showed that to be faster than raw copies on this "
                              "machine); %d host threads, one context each, keep calls in flight" % n_host,
                       "one_call_at_a_time": world * hf * bits * n_e2e / el1 / 1e9,
                       "matches_device_path": ok, "gpu_launches": sum(c_.launches for c_ in host_ctx) - l0}

        # ---- same call on the packed wire format (4 segments per byte; not a reference format) ----
        pstride = ((T + 3) // 4 + 15) // 16 * 16
        d_packed = ctx.pack_symbols(segs[:hf], T, packed_stride=pstride, stream=stream)
        stream.synchronize()
        h_packed = torch.empty((hf, pstride), dtype=torch.uint8).pin_memory()
        h_packed.copy_(d_packed)
        torch.cuda.synchronize()
        h_packed2 = [h_packed] + [h_packed.clone().pin_memory() for _ in range(n_host - 1)]
        for i in range(n_host):
            host_ctx[i].decode_batch_packed_host(code, h_packed2[i], bits, h_out[i])
        el = host_calls_in_flight(lambda i: host_ctx[i].decode_batch_packed_host(code, h_packed2[i], bits, h_out[i]),
                                  n_e2e, n_host)
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        out_p = torch.empty_like(out[:hf])
        ctx.decode_batch_packed(code, d_packed, bits, out=out_p, stream=stream)
        ev2.record(stream)
        for _ in range(n_e2e):
            ctx.decode_batch_packed(code, d_packed, bits, out=out_p, stream=stream)
        ev3.record(stream)
        stream.synchronize()
        line["packed_format"] = {"e2e": {"value": world * hf * bits * n_e2e / el / 1e9, "unit": "Gbit/s",
                                         "h2d_bytes_per_step": (hf - 1) * pstride + (T + 3) // 4,
                                         "d2h_bytes_per_step": hf * bits // 8,
                                         "api": "ced_decode_batch_packed_host"},
                                 "device_resident_value": world * hf * bits * n_e2e / (ev2.elapsed_time(ev3) * 1e-3) / 1e9,
                                 "matches_byte_format": bool(all(torch.equal(h.cuda(), out[:hf]) for h in h_out)
                                                             and torch.equal(out_p, out[:hf])),
                                 "note": "4 two-bit segments per byte; not the reference wire format, reported beside it"}
        del h_segs, h_out, h_packed, h_packed2

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

    if world == 1 and args.mode == "decode" and not args.no_subrecords:
        line.update(sub_records(ctx, ced, torch, stream, msgs[:launch], segs[:launch], out[:launch], bits, T, hbm_peak, peak_src,
                                line["roofline"]["peak"], sass, clocks))

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.mode == "decode":
        line["cpu_baseline"] = cpu_reference_rate(4.0)[0]
        # the reference's own driver shape: ONE 2048-bit packet per synchronous VITERBI_DECODER_HARD call
        # (speedDecode.c:18-23,79) through the drop-in host C library, next to the reference C on one core
        line["per_packet"] = per_packet_rate(1.0)
        line["small_batch"] = small_batch_rate(ctx, 1.0)
        line["cpu_baseline"]["one_core_2048_bit_packets"] = per_packet_reference_rate(1.0)

    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    for c_ in extra:
        c_.close()
    for c_ in host_ctx[1:]:
        c_.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def encode_roofline(frames, bits, T, ms_per_step, hbm_peak, peak_src):
    """the encoder is HBM-bound: message bytes in + one byte per coded segment out"""
    enc_bytes = frames * (bits // 8 + T)
    achieved = enc_bytes / (ms_per_step * 1e-3) / 1e9
    enc_traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            per = json.load(f).get("encodeBatchLutKernel_dram_bytes_per_launch_2p20_frames")
        if per:
            enc_traffic = per * frames / float(1 << 20)
    return {"bound": "hbm", "kernel": "encodeBatchLutKernel", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
            "frac": achieved / hbm_peak, "peak_source": peak_src, "algorithmic_bytes_per_launch": enc_bytes,
            "kernel_ms": ms_per_step, "traffic": enc_traffic}


def sub_records(ctx, ced, torch, stream, msgs, segs, out, bits, T, hbm_peak, peak_src, int_peak_tiops, sass, clocks):
    """The other BASELINE configs on the default line (one GPU): encoder (configs[2]), BER pipeline + one BER point
    (configs[3]) and the soft-decision decoder.  Each is device-timed with CUDA events on the
    launching stream, 3 warm-up passes, inputs larger than the L2."""
    code = ced.K7_DEFAULT
    frames = msgs.shape[0]
    rec = {}

    def timed(fn, n):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            fn()
        e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1) / n

    # ---- encode: BASELINE configs[2], 2^20 frames x 4096 bits ----
    ef = CONFIG3_FRAMES
    e_msgs = torch.empty((ef, bits // 8), dtype=torch.uint8, device="cuda")
    e_segs = torch.zeros((ef, SEG_STRIDE), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ctx.random_bytes(e_msgs, seed=159, stream=stream)
    l0 = ctx.launches
    ms = timed(lambda: ctx.encode_batch(code, e_msgs, out=e_segs, stream=stream), 10)
    # spot check against the decoder: the first 2^16 encoded frames decode back to their messages
    chk = ctx.decode_batch(code, e_segs[:frames], bits, stream=stream)
    stream.synchronize()
    rec["encode"] = {"metric": "K=7 r=1/2 convolutional encoded Gbit/s (information bits)",
                     "value": ef * bits / (ms * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms, "steps": 10,
                     "config": {"workload": "speedEncode K=7 rate-1/2 (0113/0171), 2^20 frames x 4096 bits (BASELINE configs[2])",
                                "frames": ef, "output": "1 byte per segment, rows of 4102 valid bytes at stride 4112"},
                     "roofline": encode_roofline(ef, bits, T, ms, hbm_peak, peak_src),
                     "round_trip_ok": bool(torch.equal(chk, e_msgs[:frames])), "gpu_launches": ctx.launches - l0}
    del e_msgs, e_segs, chk

    # ---- BER pipeline: encode -> BSC -> decode -> count on the device (configs[3]'s data path) ----
    counters = torch.zeros(4, dtype=torch.int64, device="cuda")
    b_segs = torch.zeros_like(segs)

    def ber_step():
        ctx.encode_batch(code, msgs, out=b_segs, stream=stream)
        ctx.bsc_channel(b_segs, T, 2, 0.0377, seed=2718, counters=counters[:2], stream=stream)
        ctx.decode_batch(code, b_segs, bits, out=out, stream=stream)
        ctx.ber_count(out, msgs, counters[2:], stream=stream)
    ms = timed(ber_step, 10)
    counters.zero_()
    torch.cuda.synchronize()
    ber_step()
    stream.synchronize()
    c = [int(v) for v in counters.cpu()]
    rec["ber"] = {"metric": "K=7 r=1/2 BER pipeline Gbit/s (encode+BSC+decode+count)",
                  "value": frames * bits / (ms * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_step": ms, "steps": 10,
                  "config": {"workload": "berTestK7 data path on the device, %d frames x %d bits, BSC p = 0.0377 (Eb/N0 5 dB)"
                                         % (frames, bits)},
                  "point": {"ebn0_db": 5.0, "channel_flips": c[0], "coded_bits": c[1], "decoded_errors": c[2],
                            "decoded_bits": c[3], "channel_ber": c[0] / max(1, c[1]), "decoded_ber": c[2] / max(1, c[3]),
                            "note": "berTestK7.c:96 expects 5.18e-4 +- 10 % at its 5.03 dB point (p = 0.03716); the full "
                                    "0-8 dB sweep with reference-identical subsets is tests/ber_sweep.py -> profiles/"}}
    del b_segs

    # ---- soft-decision decoder on the same frames through an AWGN channel (int8 soft symbols) ----
    soft = ctx.awgn_channel(segs.new_zeros(0, T) if False else _clean_segments(ctx, ced, msgs, T, stream), T, 3.0,
                            seed=2718, stream=stream)
    s_out = torch.empty_like(out)
    for _ in range(3):
        ctx.decode_batch_soft(code, soft, bits, out=s_out, stream=stream)
    stream.synchronize()
    ctx.set_profiling(True)
    fwd, tb = [], []
    for _ in range(6):
        ctx.decode_batch_soft(code, soft, bits, out=s_out, stream=stream)
        f, t = ctx.last_kernel_ms()
        fwd.append(f)
        tb.append(t)
    ctx.set_profiling(False)
    fwd_ms, tb_ms = sum(fwd) / len(fwd), sum(tb) / len(tb)
    ms = timed(lambda: ctx.decode_batch_soft(code, soft, bits, out=s_out, stream=stream), 10)
    cnt = torch.zeros(2, dtype=torch.int64, device="cuda")
    ctx.ber_count(s_out, msgs, cnt, stream=stream)
    stream.synchronize()
    algo_ops = frames * bits * INT_OPS_PER_BIT
    loop = sass.get("k7_soft_forward", {})
    ipf = loop.get("instr_per_frame_step")
    sm_mhz = (clocks.get("sm_mhz") or 1965.0)
    rec["soft"] = {"metric": "K=7 r=1/2 soft-decision Viterbi decoded Gbit/s", "value": frames * bits / (ms * 1e-3) / 1e9,
                   "unit": "Gbit/s", "ms_per_step": ms, "steps": 10, "dtype": "u16",
                   "config": {"workload": "2^16 frames x 4096 bits, int8 soft symbols (2 bytes per segment), BPSK+AWGN Eb/N0 3 dB"
                              if frames == FRAMES_PER_GPU else "%d frames x %d bits, int8 soft symbols" % (frames, bits),
                              "api": "ced_decode_batch_soft"},
                   "roofline": {"bound": "int_alu", "kernel": "k7SoftForwardKernel",
                                "achieved": algo_ops / (fwd_ms * 1e-3) / 1e12, "peak": int_peak_tiops, "unit": "Tiop/s",
                                "frac": algo_ops / (fwd_ms * 1e-3) / 1e12 / int_peak_tiops, "kernel_ms": fwd_ms,
                                "traceback_ms": tb_ms, "kernel_share_of_step": fwd_ms / (fwd_ms + tb_ms),
                                "issue": None if not ipf else {
                                    "instr_per_frame_step": ipf, "pipes": loop.get("pipes"),
                                    "ipc_per_sm_subpartition": frames * T * ipf / 32.0 / (148 * 4 * fwd_ms * 1e-3 * sm_mhz * 1e6)},
                                "traffic": None},
                   "check": {"decoded_bit_errors": int(cnt[0]), "decoded_bits": int(cnt[1]),
                             "ber": int(cnt[0]) / max(1, int(cnt[1])),
                             "note": "hard-decision decoding of the same channel output: BER 2.8e-2 at 3 dB "
                                     "(profiles/ber_sweep_r2_soft3.json: 2.09 dB soft-decision gain at BER 1e-4)"}}

    # ---- the same channel output quantised to 3 bits: byte metrics, one byte per segment (ced_decode_batch_softq) ----
    sigma_i8 = 32.0 * 10.0 ** (-3.0 / 20.0)
    syms = ctx.quantize_soft(soft, T, 0.6 * sigma_i8, sym_stride=SEG_STRIDE, stream=stream)
    for _ in range(3):
        ctx.decode_batch_softq(code, syms, bits, out=s_out, stream=stream)
    stream.synchronize()
    ctx.set_profiling(True)
    fwd, tb = [], []
    for _ in range(6):
        ctx.decode_batch_softq(code, syms, bits, out=s_out, stream=stream)
        f, t = ctx.last_kernel_ms()
        fwd.append(f)
        tb.append(t)
    ctx.set_profiling(False)
    qf_ms, qt_ms = sum(fwd) / len(fwd), sum(tb) / len(tb)
    qms = timed(lambda: ctx.decode_batch_softq(code, syms, bits, out=s_out, stream=stream), 10)
    cnt.zero_()
    ctx.ber_count(s_out, msgs, cnt, stream=stream)
    stream.synchronize()
    rec["soft"]["soft_3bit"] = {
        "metric": "K=7 r=1/2 3-bit soft-decision Viterbi decoded Gbit/s", "value": frames * bits / (qms * 1e-3) / 1e9,
        "unit": "Gbit/s", "ms_per_step": qms, "steps": 10, "dtype": "u8",
        "config": {"workload": "the same channel output quantised to 8 levels (step 0.6 sigma), one byte per segment",
                   "api": "ced_quantize_soft + ced_decode_batch_softq"},
        "roofline": {"bound": "int_alu", "kernel": "k7SoftQForwardKernel", "achieved": algo_ops / (qf_ms * 1e-3) / 1e12,
                     "peak": int_peak_tiops, "unit": "Tiop/s", "frac": algo_ops / (qf_ms * 1e-3) / 1e12 / int_peak_tiops,
                     "kernel_ms": qf_ms, "traceback_ms": qt_ms, "kernel_share_of_step": qf_ms / (qf_ms + qt_ms), "traffic": None},
        "check": {"decoded_bit_errors": int(cnt[0]), "decoded_bits": int(cnt[1]), "ber": int(cnt[0]) / max(1, int(cnt[1])),
                  "note": "profiles/ber_sweep_r2_soft3.json: 1.93 dB gain over hard decisions at BER 1e-4 (int8: 2.09 dB)"}}
    return rec


def _clean_segments(ctx, ced, msgs, T, stream):
    """noise-free coded symbols of msgs (the AWGN channel adds its own noise)"""
    import torch
    clean = torch.zeros((msgs.shape[0], SEG_STRIDE), dtype=torch.uint8, device="cuda")
    ctx.encode_batch(ced.K7_DEFAULT, msgs, out=clean, stream=stream)
    return clean


if __name__ == "__main__":
    main()
