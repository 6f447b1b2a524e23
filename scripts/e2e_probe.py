"""Where does the pipelined end-to-end step spend its time? (GPU box only; diagnostic, not a bench.)"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tpp_mlir_b200 import harness, xsmm  # noqa: E402

BF16 = 2
dev = torch.device("cuda:0")
torch.cuda.set_device(dev)


def raw_pcie(depth, steps=600, nbytes=512 * 1024, kernel=False):
    """torch streams: H2D(512K) [-> kernel stand-in] -> D2H(512K) per step, `depth` in flight."""
    streams = [torch.cuda.Stream() for _ in range(depth)]
    hin = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in range(depth)]
    hout = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in range(depth)]
    din = [torch.zeros(nbytes, dtype=torch.uint8, device=dev) for _ in range(depth)]
    for it in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for s in range(steps):
            d = s % depth
            st = streams[d]
            if s >= depth:
                st.synchronize()
            with torch.cuda.stream(st):
                din[d].copy_(hin[d], non_blocking=True)
                hout[d].copy_(din[d], non_blocking=True)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / steps
    return dt * 1e6


def mlp(depth, mode, steps=960):
    cfg = harness.MlpConfig(batch=256, layers=(1024, 1024, 1024, 1024), tiles=(256, 1024, 1024), dtype=BF16)

    def pinned(n):
        return (torch.rand(n) * 0.1).bfloat16().view(torch.int16).contiguous().pin_memory()

    h_w = [pinned(1024 * 1024) for _ in range(3)]
    h_b = [pinned(1024) for _ in range(3)]
    slots = [[pinned(256 * 1024)] + [torch.zeros(256 * 1024, dtype=torch.int16).pin_memory() for _ in range(3)]
             for _ in range(depth)]
    regs = h_w + h_b + [t for a in slots for t in a]
    for t in regs:
        xsmm.register_host(t, upload=True)
    h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 1024, 1024, 1024, 1024, 256 * 1024, 1 << 20, 4 | 64 | 128, 0, 5, 4, 1)
    loop = harness.NativeMlpLoop(cfg, [h] * 3, [(a, h_w, h_b) for a in slots])
    loop.run_e2e_pipelined(depth * 4, mode=mode)
    xsmm.sync()
    t0 = time.perf_counter()
    n = loop.run_e2e_pipelined(steps, mode=mode)
    dt = (time.perf_counter() - t0) / n
    xsmm.sync()
    for t in regs:
        xsmm.unregister_host(t)
    return dt * 1e6


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "batch":
    for depth in (1, 2, 4):
        print(f"raw H2D+D2H 512KiB, depth {depth}: {raw_pcie(depth):7.2f} us/step", flush=True)
    for mode, depth in (("async", 4), ("async", 8), ("batch2", 4), ("batch2", 8), ("batch3", 6), ("batch3", 9),
                        ("batch3", 12), ("batch4", 8), ("batch4", 12), ("batch6", 12), ("batch6", 18)):
        print(f"mlp e2e {mode:8s} depth {depth:2d}: {mlp(depth, mode, steps=1200):7.2f} us/step", flush=True)
    sys.exit(0)

if __name__ == "__main__":
    for depth in (1, 2, 4, 8):
        print(f"raw H2D+D2H 512KiB, depth {depth}: {raw_pcie(depth):7.2f} us/step", flush=True)
    for mode in ("async", "grouped", "streams"):
        for depth in (1, 2, 3, 4, 8, 16, 32):
            if mode == "streams" and depth > 4:
                continue
            print(f"mlp e2e {mode:8s} depth {depth:2d}: {mlp(depth, mode):7.2f} us/step", flush=True)
