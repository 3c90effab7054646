"""bench.py prints exactly one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"]


def _run(args, timeout):
    env = {k: v for k, v in os.environ.items() if not k.startswith("CED_WARP_FRAME")}   # bench.py runs the library as shipped
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                       timeout=timeout, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]           # ONE line, nothing else on stdout
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"], 300)
    for k in BASE_KEYS + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Gbit/s" and d["higher_is_better"] is True
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert "workload" in d["config"] and "model" not in d["config"] and d["vs_baseline"] is None
    assert 0.001 < d["value"] < 100.0


@pytest.mark.gpu
def test_our_arm_line():
    d = _run(["--frames", "8192", "--steps", "4", "--warmup", "3"], 900)
    for k in BASE_KEYS + ["roofline", "cpu_baseline", "clocks", "single_stream", "check"]:
        assert k in d, k
    assert d["dtype"] == "u8" and d["data"] == "synthetic" and d["scaling"] == "weak" and d["vs_baseline"] is None
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert rf["bound"] == "int_alu" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert 0.0 < rf["issue"]["ipc_per_sm_subpartition"] <= 1.0
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 8192 * 4102 - 1 and e["d2h_bytes_per_step"] == 8192 * 512 and e["value"] > 0
    assert e["matches_device_path"] is True and d["packed_format"]["matches_byte_format"] is True
    assert d["gpu_launches"] >= 2 * 4 and d["clocks"]["sm_mhz"]
    assert d["check"]["decoded_bits"] == 8192 * 4096 and 1e-4 < d["check"]["ber"] < 2e-3
    assert d["cpu_baseline"]["kind"] in ("reference", "port")
    assert e["h2d_ceiling_gbs"] > 1.0 and e["frac_of_ceiling"] > 0.0
    assert rf["issue"]["instr_per_frame_step"] == rf["issue"]["loop_instructions"] / 6.0   # read from the built library
    # the other BASELINE configs ride on the same line
    enc, ber, soft = d["encode"], d["ber"], d["soft"]
    q3 = soft["soft_3bit"]
    assert q3["value"] > soft["value"] and q3["check"]["decoded_bits"] == 8192 * 4096 and q3["check"]["ber"] < 5e-3
    assert enc["config"]["frames"] == 1 << 20 and enc["round_trip_ok"] is True and enc["roofline"]["bound"] == "hbm"
    assert 0.1 < enc["roofline"]["frac"] < 1.2 and enc["value"] > 100.0
    assert ber["point"]["decoded_bits"] == 8192 * 4096 and 1e-4 < ber["point"]["decoded_ber"] < 2e-3
    assert abs(ber["point"]["channel_ber"] - 0.0377) < 0.001
    assert soft["roofline"]["kernel"] == "k7SoftForwardKernel" and soft["value"] > 1.0 and soft["check"]["ber"] < 2e-3
    pp = d["per_packet"]                       # one 2048-bit packet per call: the two frame-parallel kernels each time
    assert pp["round_trip_ok"] is True and pp["gpu_launches"] == 2 * (pp["calls"] + 64) and pp["value"] > 0
    sb = d["small_batch"]                      # 16 such packets per call: the block + join kernels of warp_split.cu each time
    assert sb["round_trip_ok"] is True and sb["gpu_launches"] == 2 * (sb["calls"] + 10) and sb["value"] > pp["value"]
