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


@pytest.mark.parametrize("mode", ["two_stage", "fused"])
def test_uncompressed_ziq_input(built, tmp_path, mode):
    """baseband_format ziq, uncompressed (docs/pages/ZIQ.md, src-core/common/ziq.cpp:115-153,258-303): a 22-byte header + annotation in front of
    the cs16 samples; same CADUs as the raw file."""
    import struct
    O = oracle()
    cfg, raw, _ = signal("metop_ahrpt", 22)
    note = b'{"recorder": "test"}'
    inp = tmp_path / "metop.ziq"
    inp.write_bytes(b"ZIQ_" + bytes([0, 16]) + struct.pack("<Q", 6000000) + struct.pack("<Q", len(note)) + note + raw.tobytes())
    want = oracle_fec(O, cfg).run(oracle_demod(O, cfg).run(raw, stages=False)["soft"])["cadu"]
    hint = str(tmp_path / f"ziq_{mode}")
    cmd = [TOOL, "metop_ahrpt", "baseband", str(inp), hint, "--samplerate", "6e6", "--baseband_format", "ziq"] + (["--fused"] if mode == "fused" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(hint + ".cadu", np.uint8)
    assert got.size == want.size and np.array_equal(got, want)


@pytest.mark.parametrize("mode", ["two_stage", "fused"])
def test_pm_demod_cli_writes_reference_cadu_file(built, tmp_path, mode):
    """module id pm_demod (PMDemodModule's parameter set) -> ccsds_conv_concat_decoder through the host module layer."""
    O = oracle()
    cfg, raw, _ = signal("pm_bpsk", 22)
    inp = tmp_path / "pm.cs16"
    raw.tofile(inp)
    want = oracle_fec(O, cfg).run(oracle_demod(O, cfg).run(raw, stages=False)["soft"])["cadu"]
    hint = str(tmp_path / f"pm_{mode}")
    # (the oracle configuration of the test signal carries psk_demod's clock-recovery gains; pm_demod's own defaults are 0.01 / 0.01^2/4)
    gains = ["--clock_gain_mu", "8.7e-3", "--clock_gain_omega", repr(float(np.float32(pow(8.7e-3, 2) / 4.0)))]
    cmd = [TOOL, "pm_bpsk", "baseband", str(inp), hint, "--samplerate", "3e6", "--baseband_format", "cs16"] + gains + (["--fused"] if mode == "fused" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(hint + ".cadu", np.uint8)
    assert want.size >= 20 * 1024 and got.size == want.size and np.array_equal(got, want)


def test_cli_rejects_unsupported_options_loudly(built, tmp_path):
    inp = tmp_path / "x.cs16"
    np.zeros(4096, np.int16).tofile(inp)
    for extra in (["--enable_doppler", "true"], ["--baseband_format", "cu8"], ["--baseband_format", "w16"]):  # (60e6 and freq_shift used to be
        # here: the power-of-two decimator / the rotator take them now)
        base = ["--samplerate", "6e6", "--baseband_format", "cs16"]
        r = subprocess.run([TOOL, "metop_ahrpt", "baseband", str(inp), str(tmp_path / "o")] + base + extra, capture_output=True, text=True, timeout=120)
        assert r.returncode == 1 and "error" in r.stderr.lower(), (extra, r.stderr)


def test_metop_recorded_at_12msps_goes_through_the_front_end_resampler(built, tmp_path):
    """`--samplerate 12e6`: 5.14 samples/symbol > MAX_SPS, so BaseDemodModule::initb resamples to 8 MS/s (2/3) first
    (module_demod_base.cpp:66-80,203-204). Same CLI, same .cadu bytes as the reference's code."""
    import dataclasses
    import torch
    from satdump_b200 import synth
    O = oracle()
    cfg = dataclasses.replace(synth.CONFIGS["metop_ahrpt"], samplerate=12e6)
    raw, _ = synth.make_signal(cfg, 1 << 22, seed=2, device="cuda" if torch.cuda.is_available() else "cpu")
    raw = raw.cpu().numpy()
    od = oracle_demod(O, cfg)
    assert od.cfg.final_samplerate == 8e6
    want = oracle_fec(O, cfg).run(od.run(raw, stages=False)["soft"])["cadu"]
    assert want.size >= 20 * 1024
    inp = tmp_path / "metop12.cs16"
    raw.tofile(inp)
    hint = str(tmp_path / "out12")
    r = subprocess.run([TOOL, "metop_ahrpt", "baseband", str(inp), hint, "--samplerate", "12e6", "--baseband_format", "cs16"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(hint + ".cadu", np.uint8)
    assert got.size == want.size and np.array_equal(got, want)


@pytest.mark.parametrize("mode", ["two_stage", "fused"])
def test_simple_psk_decoder_cli(built, tmp_path, mode):
    """psk_demod -> ccsds_simple_psk_decoder (uncoded BPSK + NRZ-M + RS I=4) through the C++ module layer: same .cadu bytes as the
    reference's code."""
    O = oracle()
    cfg, raw, _ = signal("bpsk_simple", 22)
    inp = tmp_path / "b.cs16"
    raw.tofile(inp)
    want = oracle_fec(O, cfg).run(oracle_demod(O, cfg).run(raw, stages=False)["soft"])["cadu"]
    assert want.size >= 100 * 1024
    hint = str(tmp_path / f"simple_{mode}")
    cmd = [TOOL, "simple_bpsk", "baseband", str(inp), hint, "--samplerate", "3e6", "--baseband_format", "cs16"] + (["--fused"] if mode == "fused" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    got = np.fromfile(hint + ".cadu", np.uint8)
    assert got.size == want.size and np.array_equal(got, want)
