"""What bounds the end-to-end leg of bench.py? The same group pipeline (3 groups of 74 steps, one 37 MiB upload and one
37 MiB download per group through xsmm_cuda_upload_async / download_async / wait_host) with and without the kernel
launch in between.   python scripts/e2e_copy_probe.py"""
import sys
import time

import torch

sys.path.insert(0, ".")
from tpp_mlir_b200 import harness, xsmm

G, NG = 74, int(sys.argv[1]) if len(sys.argv) > 1 else 3
nbytes = 256 * 1024 * 2
cfg = harness.MlpConfig(batch=256, layers=(1024, 1024, 1024, 1024), tiles=(256, 1024, 1024))
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
xsmm.set_stream(torch.cuda.current_stream().cuda_stream)
w = [torch.zeros(1024 * 1024, dtype=torch.int16).pin_memory() for _ in range(3)]
b = [torch.zeros(1024, dtype=torch.int16).pin_memory() for _ in range(3)]
for t in w + b:
    xsmm.register_host(t, upload=True)
blocks, slots = [], []
for _ in range(NG):
    lvl = [torch.zeros(G, 256 * 1024, dtype=torch.int16).pin_memory() for _ in range(4)]
    for t in lvl:
        xsmm.register_host(t, upload=True)
    blocks.append(lvl)
    slots += [[x[j] for x in lvl] for j in range(G)]
rp = harness.MlpReplay(cfg, w, b, slots[0])
loop = harness.NativeMlpLoop(cfg, rp.handles, [(a, w, b) for a in slots])
LIB = xsmm.LIB


def copies_only(iters):
    for it in range(iters):
        lvl = blocks[it % NG]
        if it >= NG:
            LIB.xsmm_cuda_wait_host(lvl[3].data_ptr())
        LIB.xsmm_cuda_upload_async(lvl[0].data_ptr(), G * nbytes)
        LIB.xsmm_cuda_download_async(lvl[3].data_ptr(), G * nbytes)
    LIB.xsmm_cuda_stream_sync()


for name, fn in (("copies only", lambda n: copies_only(n // G)),
                 ("copies + kernel", lambda n: loop.run_e2e_pipelined(n, mode=f"batch{G}"))):
    fn(NG * G)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        fn(8 * NG * G)
        torch.cuda.synchronize()
        best = min(best, (time.perf_counter() - t0) / (8 * NG * G))
    print(f"{name:16s}: {best * 1e6:6.2f} us per step ({2 * nbytes / best / 1e9:.1f} GB/s both directions together)", flush=True)
