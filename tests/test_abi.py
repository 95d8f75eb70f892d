"""CPU: the C-ABI library loads, exports every symbol include/lora_b200.h declares, builds its
tables exactly like the reference constructor (bit-identical to the oracle's), and refuses to run
without a GPU instead of falling back to a CPU path."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, have_gpu
import gr_lora_b200
from gr_lora_b200 import _native as N


def test_header_symbols_are_exported():
    hdr = (ROOT / "include" / "lora_b200.h").read_text()
    declared = set(re.findall(r"\b(lora_b200_[a-z0-9_]+)\s*\(", hdr)) - {"lora_b200_frame_cb"}
    L = N.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, f"symbols declared in include/lora_b200.h but not exported: {missing}"
    assert declared == set(N.SIGNATURES), "gr_lora_b200/_native.py signatures out of sync with the header"
    assert L.lora_b200_abi_version() == 1


def test_struct_layout_matches_header():
    assert C.sizeof(N.Config) == 36 and N.Config.n_streams.offset == 16 and N.Config.trace_capacity.offset == 32
    assert C.sizeof(N.Step) == 20


@pytest.mark.parametrize("sf", range(7, 13))
def test_tables_bit_identical_to_oracle(oracle, sf):
    """build_ideal_chirps (lib/decoder_impl.cc:141-175): float phase, sincosf, ifreq of 1 and 3 chirps."""
    blob = gr_lora_b200.tables_build_host(sf=sf)
    d = oracle.Decoder(sf=sf)
    t = gr_lora_b200.split_tables(blob, d.sps)
    for name in ("downchirp", "upchirp", "downchirp_ifreq", "upchirp_ifreq", "upchirp_ifreq_v"):
        assert np.array_equal(t[name].view(np.uint8), getattr(d, name).view(np.uint8)), name
    tw = t["twiddles"]
    j = np.arange(d.sps)
    np.testing.assert_allclose(tw, np.exp(-2j * np.pi * j / d.sps), atol=1e-7)


def test_invalid_sf_is_rejected_like_the_reference(capsys):
    """lib/decoder_impl.cc:57-61: message on stderr and exit(1)."""
    with pytest.raises(SystemExit) as e:
        gr_lora_b200.decoder(1e6, 125000, 5, False, 4, True)
    assert e.value.code == 1
    assert "Spreading factor should be between 6 and 12" in capsys.readouterr().err


@pytest.mark.skipif(have_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CUDA device"):
        gr_lora_b200.decoder(1e6, 125000, 7, False, 4, True)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under gr_lora_b200/ or include/ may reference it."""
    bad = []
    for p in list((ROOT / "gr_lora_b200").rglob("*")) + list((ROOT / "include").rglob("*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".h", ".cc", ".cpp"):
            txt = p.read_text(errors="ignore")
            if re.search(r"^\s*(from|import)\s+oracle\b|lora_oracle|liblora_oracle|oracle/", txt, re.M):
                bad.append(str(p))
    assert not bad, bad


def test_coding_rate_above_4_is_rejected():
    """The reference aborts in deinterleave for more than 8 bits per word (lib/decoder_impl.cc:541-545); the library
    refuses the configuration up front (device-independent check) instead of overrunning its words array."""
    for cr in (5, 6, 7):
        with pytest.raises(RuntimeError, match="coding rate must be 0..4"):
            gr_lora_b200.decoder(1e6, 125000, 7, True, cr, True)


def test_documented_defaults_match_the_implementation():
    hdr = (ROOT / "include" / "lora_b200.h").read_text()
    src = (ROOT / "gr_lora_b200" / "csrc" / "lora_b200.cu").read_text()
    assert "max_frames_per_call;/* per stream (0 = 8)" in hdr
    assert "if (d->cfg.max_frames_per_call == 0) d->cfg.max_frames_per_call = 8;" in src


def test_every_abi_entry_point_is_mapped_in_integration_md():
    """INTEGRATION.md is the map from the reference's interface to the C ABI: every function include/lora_b200.h declares
    must have a row there."""
    import re
    root = Path(__file__).resolve().parent.parent
    names = sorted(set(re.findall(r"\b(lora_b200_[a-z0-9_]+)\s*\(", (root / "include" / "lora_b200.h").read_text())))
    doc = (root / "INTEGRATION.md").read_text()
    missing = [n for n in names if n not in doc]
    assert len(names) > 40 and not missing, missing
