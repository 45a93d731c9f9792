"""Copy / read-only / write-only HBM bandwidth with library kernels (context for the roofline
fractions of write-dominated kernels: batch_filter and the smoothers write ~95 % of their bytes)."""
import json
import numpy as np
import torch

n = 1 << 28                                    # 1 GiB of fp32
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


out = {}
ms = timeit(lambda: b.copy_(a)); out["copy_GBps"] = 2 * n * 4 / ms / 1e6
ms = timeit(lambda: a.fill_(1.0)); out["write_only_GBps"] = n * 4 / ms / 1e6
ms = timeit(lambda: a.sum()); out["read_only_GBps"] = n * 4 / ms / 1e6
print(json.dumps(out))
