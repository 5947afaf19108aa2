"""What a pure write stream reaches on this GPU: cudaMemset / torch fill of a one-hot-sized buffer (17.2 GB), next to the
device-to-device copy MEASURED_PEAKS.json quotes.  The fused step kernel is 97 % writes, so this is its ceiling.

    python tools/write_ceiling.py
"""
import json

import torch

n = (1 << 24) * 1024  # bytes of the fp32 one-hot of 2^24 games
buf = torch.empty(n, dtype=torch.uint8, device="cuda")
src = torch.empty(n // 2, dtype=torch.uint8, device="cuda")
res = {}


def timed(name, fn, nbytes, reps=10):
    for _ in range(2):
        fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    res[name] = {"ms": best, "GB/s": nbytes / best / 1e6}


timed("memset_u8 (cudaMemsetAsync, 17.2 GB)", lambda: buf.zero_(), n)
timed("fill_f32 (elementwise kernel, 17.2 GB)", lambda: buf.view(torch.float32).fill_(1.0), n)
timed("copy d2d (8.6 GB read + 8.6 GB written)", lambda: buf[: n // 2].copy_(src), n)
timed("read (sum of 17.2 GB as int32)", lambda: buf.view(torch.int32).sum(), n, reps=5)
print(json.dumps(res, indent=1))
