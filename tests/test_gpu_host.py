"""GPU: the C++ host module layer (satdump_b200/host, the code the SatDump plugin shim wraps) run as the stand-alone
`b200_pipeline` tool, the way `satdump pipeline metop_ahrpt baseband in.cs16 out --samplerate 6e6 --baseband_format cs16` is run:
two modules joined by a byte FIFO, and the fused single-module variant. Outputs are the reference's file formats."""
import os
import subprocess

import numpy as np
import pytest

from tests.common import ROOT, oracle, oracle_demod, oracle_fec, signal

pytestmark = pytest.mark.gpu
TOOL = os.path.join(ROOT, "satdump_b200", "host", "b200_pipeline")


@pytest.mark.parametrize("mode", ["two_stage", "fused"])
def test_metop_cli_writes_reference_cadu_file(built, tmp_path, mode):
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 22)
    inp = tmp_path / "metop.cs16"
    raw.tofile(inp)
    want = oracle_fec(O, cfg).run(oracle_demod(O, cfg).run(raw, stages=False)["soft"])["cadu"]
    hint = str(tmp_path / f"out_{mode}")
    cmd = [TOOL, "metop_ahrpt", "baseband", str(inp), hint, "--samplerate", "6e6", "--baseband_format", "cs16"] + (["--fused"] if mode == "fused" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(hint + ".cadu", np.uint8)
    # the demodulator is fed in batches of its own choosing: same stream, same frames
    assert got.size == want.size and np.array_equal(got, want)


def test_cli_rejects_unsupported_options_loudly(built, tmp_path):
    inp = tmp_path / "x.cs16"
    np.zeros(4096, np.int16).tofile(inp)
    for extra in (["--dc_block", "true"], ["--baseband_format", "cu8"], ["--samplerate", "60e6"]):
        base = ["--samplerate", "6e6", "--baseband_format", "cs16"]
        r = subprocess.run([TOOL, "metop_ahrpt", "baseband", str(inp), str(tmp_path / "o")] + base + extra, capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "error" in r.stderr.lower(), (extra, r.stderr)
