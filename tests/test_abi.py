"""The C-ABI shared library loads on a CPU-only box, exports every symbol include/bke.h declares,
validates arguments, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from filterpy_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "bke.h")).read()
    declared = set(re.findall(r"\b(bke_[a-z0-9_]+)\s*\(", hdr))
    declared = {d for d in declared if not d.endswith("_args")}
    assert declared == set(_lib.EXPORTED_SYMBOLS), declared ^ set(_lib.EXPORTED_SYMBOLS)
    for s in declared:
        assert getattr(lib, s) is not None
    assert lib.bke_abi_version() == 1


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors must have the layout a C compiler gives the structs of include/bke.h:
    a probe compiled with gcc prints sizeof / offsetof of every struct."""
    import subprocess
    structs = {"bke_kf_args": _lib.KfArgs, "bke_kf_batch_args": _lib.KfBatchArgs, "bke_ukf_args": _lib.UkfArgs,
               "bke_resample_shard_args": _lib.ResampleShardArgs, "bke_resample_shard_ext": _lib.ResampleShardExt, "bke_rts_args": _lib.RtsArgs, "bke_ukf_rts_args": _lib.UkfRtsArgs,
               "bke_mm_args": _lib.MmArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "bke.h"', 'int main(void) {']
    for cname, cls in structs.items():
        lines.append('printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ['return 0; }']
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split("\n")
    seen = 0
    for ln in out:
        if not ln.strip():
            continue
        cname, fname, val = ln.split()
        cls = structs[cname]
        if fname == "sizeof":
            assert ctypes.sizeof(cls) == int(val), (cname, ctypes.sizeof(cls), val)
        else:
            assert getattr(cls, fname).offset == int(val), (cname, fname)
        seen += 1
    assert seen == sum(len(c._fields_) + 1 for c in structs.values())
    assert _lib.KfArgs.alpha_sq.offset == 32 and _lib.KfArgs.x.offset == 40


def test_argument_validation_without_gpu():
    lib = _lib.load()
    a = _lib.KfArgs()
    a.n_filters, a.dim_x, a.dim_z, a.dtype, a.flags = 4, 0, 1, _lib.BKE_F64, 3
    assert lib.bke_kf_step(a, None) == _lib.BKE_ERR_BAD_ARG
    assert b"dim_x must be 1 or greater" in lib.bke_last_error()
    a.dim_x = 2
    a.flags = 0
    assert lib.bke_kf_step(a, None) == _lib.BKE_ERR_BAD_ARG
    with pytest.raises(ValueError):
        _lib.check(_lib.BKE_ERR_BAD_ARG)
    assert lib.bke_resample_workspace_bytes(1 << 20) > (1 << 20) // 2048 * 8
    assert lib.bke_systematic_resample(-1, None, 0.5, None, None, 0, None, None, None) == _lib.BKE_ERR_BAD_ARG


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    assert lib.bke_device_count() == 0
    x = np.zeros((1, 2)); P = np.eye(2)[None].copy(); F = np.eye(2); Q = np.eye(2)
    a = _lib.KfArgs()
    a.n_filters, a.dim_x, a.dim_z, a.dtype, a.flags, a.alpha_sq = 1, 2, 1, _lib.BKE_F64, _lib.BKE_DO_PREDICT, 1.0
    a.x = a.x_out = x.ctypes.data; a.P = a.P_out = P.ctypes.data; a.F = F.ctypes.data; a.Q = Q.ctypes.data
    assert lib.bke_kf_step(a, None) == _lib.BKE_ERR_CUDA
    assert b"no CPU fallback" in lib.bke_last_error()
    from filterpy_b200.kalman import KalmanFilter
    from filterpy_b200.monte_carlo import systematic_resample
    with pytest.raises(_lib.BkeError):
        KalmanFilter(2, 1)
    with pytest.raises(_lib.BkeError):
        systematic_resample([.5, .5])
    from filterpy_b200.monte_carlo import residual_resample, multinomial_resample
    for fn in (residual_resample, multinomial_resample):
        with pytest.raises(_lib.BkeError):
            fn([.5, .5])
    # the filter whose steps run on the tensor-core tile has no CPU path either
    with pytest.raises(_lib.BkeError):
        KalmanFilter(16, 4, n_filters=8, dtype=np.float32)


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (only tests/, smoke() and bench.py may)."""
    pkg = os.path.join(ROOT, "filterpy_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src, f
