"""roofline.issue in bench.py is derived from the SHIPPED binary: convolutionalencdec_b200/sass_stats.json must
describe the libced_cuda.so that is in the tree (tools/sass_loop_stats.py, run by `make cuda`), and the loop
listings under profiles/ must be the same loops."""
import json
import os
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT


@pytest.mark.skipif(shutil.which("cuobjdump") is None, reason="cuobjdump not on PATH")
def test_sass_stats_match_the_built_library(tmp_path):
    out = tmp_path / "stats.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_loop_stats.py"), "--out", str(out),
                    "--sass-dir", str(tmp_path)], check=True, stdout=subprocess.DEVNULL)
    fresh = json.load(open(out))
    shipped = json.load(open(os.path.join(ROOT, "convolutionalencdec_b200", "sass_stats.json")))
    assert fresh == shipped
    fwd = fresh["k7_forward"]
    assert fwd["steps_per_iteration"] == 6 and 600 < fwd["loop_instructions"] < 800
    assert fwd["pipes"]["alu"] + fwd["pipes"]["fma"] > 0.9 * fwd["loop_instructions"]
    assert "STG" in fwd["opcodes"] and "PRMT" in fwd["opcodes"]
    listing = open(os.path.join(ROOT, "profiles", "k7_forward_loop.sass")).read().splitlines()
    assert len([l for l in listing if l.startswith("/*")]) == fwd["loop_instructions"]
    assert 1200 < fresh["k7_soft_forward"]["loop_instructions"] < 1600
