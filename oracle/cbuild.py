"""Build + load ``oracle/_build/liboracle.so`` (gcc; TEST INFRASTRUCTURE, see oracle/__init__.py)."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "_build/liboracle.so"])
    return _SO


def load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_systematic_resample.restype = ctypes.c_int
        _lib.oracle_stratified_resample.restype = ctypes.c_int
        _lib.oracle_kf_step_f64.restype = ctypes.c_int
    return _lib
