"""GPU: the compiled C++ drop-in (gr_lora_b200/host/decoder_impl.cc = the body a gr-lora maintainer would
give gr::lora::decoder_impl) run under a fake GNU Radio scheduler on a cf32 file, the way
apps/lora_receive_file_nogui.py runs the reference block.  std::cout must read like README.md:77-85."""
import json
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import FRAME_CASES, case_decoder_args, make_case_iq

pytestmark = pytest.mark.gpu
GOLD = json.loads((Path(__file__).parent / "golden" / "golden.json").read_text())


@pytest.fixture(scope="module")
def shim():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    from gr_lora_b200 import build as B
    return B.build_shim()


def run_shim(shim, x, case, tmp_path, chunk=None, env=None):
    f = tmp_path / "capture.cf32"
    np.ascontiguousarray(x, np.complex64).tofile(f)
    a = case_decoder_args(case)
    cmd = [str(shim), str(f), "1000000", "125000", str(a["sf"]), str(int(a["implicit"])), str(a["cr"]), str(int(a["crc"])),
           str(int(a["reduced_rate"]))] + ([str(chunk)] if chunk else [])
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env={**os.environ, **(env or {})})
    assert p.returncode == 0, p.stderr
    frames = [ln.split()[1] for ln in p.stderr.splitlines() if ln.startswith("FRAME ")]
    consumed = [int(ln.split()[1]) for ln in p.stderr.splitlines() if ln.startswith("CONSUMED ")][0]
    return p.stdout, frames, consumed


def test_readme_console_output(shim, tmp_path):
    from gr_lora_b200 import tx
    fs = tx.encode_frame(bytes.fromhex("deadbeef700d"), 7, 4)
    x = tx.channel([tx.modulate_frame(fs, 7)] * 5, sf=7, snr_db=40.0, seed=0x4C6F5201, gap_symbols=97.66)
    out, frames, _ = run_shim(shim, x, FRAME_CASES[0], tmp_path)
    assert out.startswith(GOLD["readme"]["banner"])
    lines = out[len(GOLD["readme"]["banner"]):].splitlines()
    assert len(lines) == 5 and all(ln.startswith(GOLD["readme"]["line"]) for ln in lines)
    assert [f[30:] for f in frames] == ["049040deadbeef700d"] * 5


@pytest.mark.parametrize("name", ["sf8_cr4", "sf9_cr3", "sf10_cr1_implicit", "sf11_cr4_rr"])
def test_frames_equal_fixture_through_cpp_block(shim, tmp_path, name):
    case = [c for c in FRAME_CASES if c[0] == name][0]
    x, _, _ = make_case_iq(case)
    g = GOLD["frames"][name]
    for chunk in (None, 8 * (8 << case[1])):        # one big work() call / GNU-Radio-sized calls of 8 symbols
        out, frames, consumed = run_shim(shim, x, case, tmp_path, chunk)
        assert frames == g["frames"]
        assert consumed == g["consumed"]
        assert out == g["stdout"]                   # banner + hex lines, byte for byte what the oracle "printed"


def test_bad_sf_exits_like_the_reference(shim, tmp_path):
    f = tmp_path / "empty.cf32"
    np.zeros(16, np.complex64).tofile(f)
    p = subprocess.run([str(shim), str(f), "1000000", "125000", "5", "0", "4", "1"], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "Spreading factor should be between 6 and 12" in p.stderr
