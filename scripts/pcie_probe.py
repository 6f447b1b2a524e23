"""Host<->device copy rates of this box (pinned memory), alone and in both directions at once: the floor of the
end-to-end step (512 KiB up + 512 KiB down per forward pass).   python scripts/pcie_probe.py"""
import time

import torch

dev = torch.device("cuda", 0)
for chunk_kib, n in ((512, 296), (37 * 1024, 4)):
    nbytes = chunk_kib * 1024
    h_in = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n)]
    h_out = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(n)]
    d_in = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(n)]
    d_out = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(n)]
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down, reps=5):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            for i in range(n):
                if up:
                    with torch.cuda.stream(s_up):
                        d_in[i].copy_(h_in[i], non_blocking=True)
                if down:
                    with torch.cuda.stream(s_down):
                        h_out[i].copy_(d_out[i], non_blocking=True)
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        return nbytes * n * reps / dt / 1e9, t_issue / (reps * n * (int(up) + int(down))) * 1e6

    run(True, True, 1)
    for name, up, down in (("H2D only", True, False), ("D2H only", False, True), ("both directions", True, True)):
        gbs, issue_us = run(up, down)
        print(f"{chunk_kib:6d} KiB chunks, {name:16s}: {gbs:6.1f} GB/s per direction, host issue {issue_us:.2f} us per copy",
              flush=True)
