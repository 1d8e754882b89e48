"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/liboracle.so, our own C restatement (oracle.c).
Same Python API as oracle.ref (Demod, Fec, cc_decode, ...), minus the reference-only helpers."""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("oracle._port_impl", os.path.join(_HERE, "ref.py"))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
from . import ref as _r  # share the ctypes struct classes so cfg objects are interchangeable

_m.DemodCfg, _m.FecCfg = _r.DemodCfg, _r.FecCfg
_m._PATH = os.path.join(_HERE, "liboracle.so")
_m._PFX = "orc_"

_orig_lib = _m.lib


def _lib():
    import ctypes as C
    if _m._lib is None:
        L = _m._Prefixed(C.CDLL(_m._PATH), _m._PFX)
        L.ref_demod_create.restype = C.c_void_p
        L.ref_demod_create.argtypes = [C.POINTER(_m.DemodCfg)]
        L.ref_demod_destroy.argtypes = [C.c_void_p]
        L.ref_demod_sps.argtypes = [C.c_void_p]
        L.ref_demod_sps.restype = C.c_float
        L.ref_demod_run.restype = C.c_long
        L.ref_demod_run.argtypes = [C.c_void_p, C.c_void_p, C.c_long] + [C.c_void_p] * 5 + [C.c_long]
        L.ref_demod_state.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_demod_run_stage.restype = C.c_long
        L.ref_demod_run_stage.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.ref_demod_last_front.restype = C.c_long
        L.ref_demod_last_front.argtypes = [C.c_void_p]
        L.ref_demod_pm_dumps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_demod_pm_state.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_fast_atan2f.restype = C.c_float
        L.ref_fast_atan2f.argtypes = [C.c_float, C.c_float]
        L.ref_fast_cos.restype = C.c_float
        L.ref_fast_cos.argtypes = [C.c_float]
        L.ref_fast_sin.restype = C.c_float
        L.ref_fast_sin.argtypes = [C.c_float]
        L.ref_resample.restype = C.c_long
        L.ref_resample.argtypes = [C.POINTER(_m.DemodCfg), C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        L.ref_resampler_taps.argtypes = [C.c_uint, C.c_uint, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ref_fec_create.restype = C.c_void_p
        L.ref_fec_create.argtypes = [C.POINTER(_m.FecCfg)]
        L.ref_fec_destroy.argtypes = [C.c_void_p]
        L.ref_fec_chunk_size.argtypes = [C.c_void_p]
        L.ref_fec_cadu_bytes.argtypes = [C.c_void_p]
        L.ref_fec_run.restype = C.c_long
        L.ref_fec_run.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_void_p, C.c_long] + [C.c_void_p] * 7
        L.ref_rs_decode_interleaved.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_derand.argtypes = [C.c_void_p, C.c_int]
        L.ref_cc_encode.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_cc_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.ref_rotate_soft.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.ref_deframe.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.ref_rrc_taps.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]
        L.ref_mm_bank.argtypes = [C.c_void_p]
        L.ref_demux_create.restype = C.c_void_p
        L.ref_demux_create.argtypes = [C.c_int] * 4
        L.ref_demux_destroy.argtypes = [C.c_void_p]
        L.ref_demux_run.restype = C.c_long
        L.ref_demux_run.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_int, C.c_ulonglong, C.c_long, C.c_void_p, C.c_long, C.POINTER(C.c_long), C.c_void_p, C.c_long]
        L.ref_pipeline_run.restype = C.c_long
        L.ref_pipeline_run.argtypes = [C.POINTER(_m.DemodCfg), C.POINTER(_m.FecCfg), C.c_void_p, C.c_long, C.c_void_p, C.c_long]
        _m._lib = L
    return _m._lib


_m.lib = _lib


def available():
    return os.path.exists(_m._PATH)


lib = _lib
DemodCfg, FecCfg = _m.DemodCfg, _m.FecCfg
demod_cfg, metop_cfg, ccsds_cfg, simple_cfg = _m.demod_cfg, _m.metop_cfg, _m.ccsds_cfg, _m.simple_cfg
final_samplerate_of, resample, resampler_taps = _m.final_samplerate_of, _m.resample, _m.resampler_taps
Fec = _m.Fec
Demux = _m.Demux
run_stage = _m.run_stage
rs_decode_interleaved, derand, cc_encode, cc_decode, rotate_soft, deframe = (
    _m.rs_decode_interleaved, _m.derand, _m.cc_encode, _m.cc_decode, _m.rotate_soft, _m.deframe)


class Demod(_m.Demod):
    @property
    def buffer_size(self):
        fs = int(self.cfg.samplerate)
        return self.cfg.buffer_size or min(1000000, max(8193, fs // 200))


def rrc_design(gain, fs, rs, alpha, ntaps):
    import numpy as np
    out = np.zeros((ntaps | 1) + 2, np.float32)
    n = _lib().ref_rrc_taps(gain, fs, rs, alpha, ntaps, out.ctypes.data)
    return out[:n].copy()


def mm_taps():
    import numpy as np
    out = np.zeros(128 * 8, np.float32)
    _lib().ref_mm_bank(out.ctypes.data)
    return out.reshape(128, 8)


def pipeline_run(dcfg, fcfg, raw):
    """Single-thread end-to-end port. Returns CADU bytes."""
    import numpy as np
    raw = np.ascontiguousarray(raw)
    n = raw.size if dcfg.format == 0 and np.iscomplexobj(raw) else raw.size // 2
    cap = int(n * 0.2) + 65536
    cadu = np.zeros(cap, np.uint8)
    w = _lib().ref_pipeline_run(dcfg, fcfg, raw.ctypes.data, n, cadu.ctypes.data, cap)
    return cadu[:w].copy()


def agc_exact(x, rate=1e-2, ref=1.0, max_gain=65536.0):
    """AGCBlock's recurrence (agc.cpp:25-39) evaluated in double precision on complex64 input: the reference's float output deviates
    from this by its own accumulated rounding noise, which is the floor of the AGC parity gate."""
    import ctypes as C
    import numpy as np
    L = C.CDLL(_m._PATH)
    L.orc_agc_exact.argtypes = [C.c_void_p, C.c_long, C.c_double, C.c_double, C.c_double, C.c_void_p]
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty_like(x)
    L.orc_agc_exact(x.ctypes.data, x.size, float(np.float32(rate)), ref, max_gain, out.ctypes.data)
    return out
