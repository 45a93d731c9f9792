"""Chain-speed probe: a shard whose running sum stays inside ONE binade (carry 0.5, weights summing to
0.4): no crossings, no wrong binade guesses — what the look-back pipeline can do when nothing is rare."""
import ctypes, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from filterpy_b200 import _lib
from filterpy_b200.common import workloads as wl
from filterpy_b200._dev import stream_ptr
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 26
N = 1 << lg
lib = _lib.load()
w = torch.from_numpy(wl.resample_weights(N, "heavy", seed=97) * 0.4).cuda()
carry = torch.tensor([0.5], dtype=torch.float64, device="cuda")
ws_bytes = int(lib.bke_resample_workspace_bytes(N))
ws = torch.empty(ws_bytes + 256, dtype=torch.uint8, device="cuda")
ws_ptr = ws.data_ptr() + (-ws.data_ptr()) % 256
idx = torch.empty(N, dtype=torch.int32, device="cuda")
info = torch.zeros(8, dtype=torch.int32, device="cuda")
rng = torch.zeros(2, dtype=torch.int64, device="cuda")
cout = torch.zeros(1, dtype=torch.float64, device="cuda")
a = _lib.ResampleShardArgs()
a.n_local = N; a.n_global = N; a.j_offset = 0; a.capacity = N
a.weights = w.data_ptr(); a.uniforms = None; a.u = 0.0763
a.carry_approx = carry.data_ptr(); a.carry_exact = carry.data_ptr()
a.indexes = idx.data_ptr(); a.out_range = rng.data_ptr(); a.carry_out = cout.data_ptr()
a.workspace = ws_ptr; a.workspace_bytes = ws_bytes; a.info = info.data_ptr(); a.is_last = 0; a.phase = 7
def run():
    _lib.check(lib.bke_resample_shard(ctypes.byref(a), stream_ptr(torch.device("cuda", 0))))
for _ in range(3):
    run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
ev[0].record()
for i in range(5):
    run(); ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
print("ONEBINADE 2^%d min=%.3f ms frac=%.3f info=%s range=%s carry_out=%.6f" % (lg, min(ms), 12.0 * N / (min(ms) * 1e-3) / 6571.6e9, info.cpu().tolist(), rng.cpu().tolist(), float(cout.item())))
