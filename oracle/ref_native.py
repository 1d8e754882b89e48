"""TEST INFRASTRUCTURE ONLY — the same reference translation units as oracle.ref, compiled the way the reference's own release build
does (-O3 -march=<AVX2+FMA> -ffp-contract=fast: FMA contraction; oracle/Makefile NATIVE_FLAGS) into oracle/_ref/libsatref_native.so.
Same Python API as oracle.ref. Used as the reference-vs-reference floor of the float parity gates and as the faster CPU baseline."""
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("oracle._ref_native_impl", os.path.join(_HERE, "ref.py"))
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
from . import ref as _r  # share the ctypes struct classes so cfg objects are interchangeable

_m.DemodCfg, _m.FecCfg = _r.DemodCfg, _r.FecCfg
_m._PATH = os.path.join(_HERE, "_ref", "libsatref_native.so")


def available():
    return os.path.exists(_m._PATH)


lib = _m.lib
DemodCfg, FecCfg = _m.DemodCfg, _m.FecCfg
demod_cfg, metop_cfg, ccsds_cfg, simple_cfg = _m.demod_cfg, _m.metop_cfg, _m.ccsds_cfg, _m.simple_cfg
Demod, Fec, run_stage, resample, pipeline_timed = _m.Demod, _m.Fec, _m.run_stage, _m.resample, _m.pipeline_timed
