"""Device-memory plumbing shared by the host-side mirrors (torch is used for allocation, streams
and host<->device copies only; all arithmetic happens in libbke.so)."""
import numpy as np
import torch

from . import _lib

_TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}


def resolve_dtype(dtype):
    if isinstance(dtype, torch.dtype):
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("dtype must be float32 or float64")
        return dtype
    return _TORCH[np.dtype(dtype)]


def bke_dtype(tdtype):
    return _lib.BKE_F32 if tdtype == torch.float32 else _lib.BKE_F64


def require_cuda(device):
    """Resolve the device; fail loudly when no GPU is present (no CPU fallback)."""
    if not torch.cuda.is_available():
        raise _lib.BkeError("filterpy_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    dev = torch.device(device if device is not None else "cuda")
    if dev.type != "cuda":
        raise ValueError("device must be a CUDA device")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


def to_dev(a, dtype, device):
    """numpy / list / scalar / torch tensor -> contiguous device tensor of `dtype`."""
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=dtype).contiguous()
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64 if dtype == torch.float64 else np.float32))
    return torch.from_numpy(arr).to(device)


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device):
    return torch.cuda.current_stream(device).cuda_stream
