"""CPU: the C++ host module layer's parameter handling (satdump_b200/host/stream_modules.cpp), through the b200_pipeline CLI. Parameters
are validated before any device is touched, so the messages can be checked without a GPU; with valid parameters and no GPU the tool
must stop with the library's ENODEV text — never decode on the host."""
import os
import subprocess

import pytest

from tests.common import ROOT

TOOL = os.path.join(ROOT, "satdump_b200", "host", "b200_pipeline")
BASE = ["--samplerate", "6e6", "--baseband_format", "cs16"]


def run(pipe, extra, tmp_path):
    inp = tmp_path / "in.cs16"
    inp.write_bytes(b"\0" * 4096)
    return subprocess.run([TOOL, pipe, "baseband", str(inp), str(tmp_path / "out")] + BASE + extra, capture_output=True, text=True, timeout=60)


@pytest.mark.parametrize("pipe,extra,needle", [
    ("pm_bpsk", ["--samplerate", "3e6", "--pll_bw", "0.9"], "pll_bw"),                         # pm_demod's carrier PLL bandwidth out of range
    ("metop_ahrpt", ["--enable_doppler", "true"], "enable_doppler"),
    ("metop_ahrpt", ["--has_carrier", "true"], "carrier mode, constellation must be BPSK"),     # module_psk_demod.cpp:95-96
    ("simple_bpsk", ["--samplerate", "3e6", "--has_carrier", "true"], "Carrier PLL Bw"),          # :98-102
    ("metop_ahrpt", ["--baseband_format", "cu8"], "baseband_format"),
    ("metop_ahrpt", ["--samplerate", "1e6"], "sampling rate is too low"),         # module_demod_base.cpp:96-105
    ("simple_qpsk", ["--symbolrate", "2400000", "--oqpsk_method2", "true"], "oqpsk_method2"),
    ("simple_bpsk", ["--samplerate", "3e6", "--hard_symbols", "true"], "hard_symbols"),
    ("simple_bpsk", ["--samplerate", "3e6", "--constellation", "8psk"], "invalid constellation"),
    ("jpss_hrd", ["--samplerate", "50e6", "--conv_rate", "4/5"], "conv_rate"),                  # (2/3 ... 7/8 are the Viterbi_Depunc rates)
])
def test_unsupported_parameters_are_named(built, tmp_path, pipe, extra, needle):
    r = run(pipe, extra, tmp_path)
    assert r.returncode == 1 and needle.lower() in r.stderr.lower(), r.stderr


def test_valid_parameters_reach_the_device_check(built, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    for pipe, extra in [("metop_ahrpt", []), ("metop_ahrpt", ["--samplerate", "12e6"]), ("metop_ahrpt", ["--dc_block", "true", "--iq_swap", "true"]),
                        ("simple_bpsk", ["--samplerate", "3e6", "--post_costas_dc", "true"]), ("metop_ahrpt", ["--freq_shift", "-100000"]),
                        ("simple_bpsk", ["--samplerate", "3e6", "--has_carrier", "true", "--carrier_pll_bw", "0.001"]),
                        ("pm_bpsk", ["--samplerate", "3e6"]), ("pm_bpsk", ["--samplerate", "6e6", "--symbolrate", "250000", "--resample_after_pll", "true"])]:
        r = run(pipe, extra, tmp_path)
        assert r.returncode == 1 and "no cpu fallback" in r.stderr.lower() and "cuda" in r.stderr.lower(), (pipe, extra, r.stderr)


def _ziq(path, payload, bits=16, compressed=0, signature=b"ZIQ_", annotation=b'{"note": "x"}'):
    import struct
    path.write_bytes(signature + bytes([compressed, bits]) + struct.pack("<Q", 6000000) + struct.pack("<Q", len(annotation)) + annotation + payload)


def test_ziq_header_is_read_before_the_device_is_touched(built, tmp_path):
    """baseband_format ziq (src-core/common/ziq.cpp:115-153): uncompressed files are accepted (their samples are cs8 / cs16 / cf32 behind the
    header), ZSTD-compressed ones and files without the signature are refused by name."""
    import torch
    out = str(tmp_path / "o")
    good, comp, bad = tmp_path / "a.ziq", tmp_path / "b.ziq", tmp_path / "c.ziq"
    _ziq(good, b"\0" * 65536)
    _ziq(comp, b"\0" * 65536, compressed=1)
    _ziq(bad, b"\0" * 65536, signature=b"RIFF")
    for fused in ([], ["--fused"]):
        args = ["--samplerate", "6e6", "--baseband_format", "ziq"] + fused
        r = subprocess.run([TOOL, "metop_ahrpt", "baseband", str(comp), out] + args, capture_output=True, text=True, timeout=60)
        assert r.returncode == 1 and "zstd" in r.stderr.lower(), r.stderr
        r = subprocess.run([TOOL, "metop_ahrpt", "baseband", str(bad), out] + args, capture_output=True, text=True, timeout=60)
        assert r.returncode == 1 and "not a valid ziq" in r.stderr.lower(), r.stderr
        if not torch.cuda.is_available():
            r = subprocess.run([TOOL, "metop_ahrpt", "baseband", str(good), out] + args, capture_output=True, text=True, timeout=60)
            assert r.returncode == 1 and "no cpu fallback" in r.stderr.lower(), r.stderr
