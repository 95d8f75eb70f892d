"""GPU parity of the opt-in K1 kernel variants (file name: runs after every other GPU test, so an experimental
variant can never mask the default paths under `pytest -x`).  The kernel choice is read from the environment once per process, so
each variant runs in its own interpreter (tools/k1_ab.py): bins bit-equal to the oracle on true symbols of that SF
(ragged count, several grid passes, edge bins), magnitudes within 1e-4, every launch under the hang watchdog."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def run_variant(sf, env):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    e = dict(os.environ)
    e.update(env)
    e["LORA_B200_XG_WATCHDOG"] = "1"
    p = subprocess.run([sys.executable, str(ROOT / "tools" / "k1_ab.py"), "--sf", str(sf), "--gib", "0.25", "--reps", "2"],
                       capture_output=True, text=True, env=e, timeout=240)
    assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
    return json.loads(p.stdout.strip().splitlines()[-1])


@pytest.mark.parametrize("sf,env", [
    (10, {"LORA_B200_K1_XCHG": "0"}),                                   # k1_xchg<10,256>: teams of 2
    (11, {"LORA_B200_K1_XCHG": "1"}),                                   # k1_xchg<11,256>: teams of 4
    (12, {"LORA_B200_K1_XCHG": "2", "LORA_B200_K1_XCHG_T": "128"}),     # k1_xchg<12,128>: teams of 16
    (12, {"LORA_B200_K1_SF12_GENERIC": "1"}),                           # the DIF-split k1_fft_kernel<12>
    (12, {"LORA_B200_K1_AB": "2", "LORA_B200_K1_AB_NA": "64"}),         # k1_ab<12>: producer / consumer roles, radix 32
    (10, {"LORA_B200_K1_AB": "0", "LORA_B200_K1_AB_NA": "64"}),         # k1_ab<10>: radix 8, 4 symbols per producer item
])
def test_k1_variant_matches_oracle(sf, env):
    out = run_variant(sf, env)
    assert out["parity"]["bins_equal"] and out["parity"]["mags_close"], out
    assert out["parity"]["vs_tx"] == 1.0
