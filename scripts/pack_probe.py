"""tensor.pack of a 4096 x 4096 bf16 matrix into 32 x 32 tiles as the reference lowers it (16384 unary identity invokes),
captured: one launch of the TMA-to-TMA grid copy (tile_grid.cu). Timing + a profiling target for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpp_mlir_b200 import harness, xsmm

M = N = 4096
sets = 5
A = [torch.randint(0, 30000, (M * N,), dtype=torch.int16, device="cuda") for _ in range(sets)]
B = [torch.zeros(M * N, dtype=torch.int16, device="cuda") for _ in range(sets)]
rp = harness.PackReplay(xsmm.BF16, M, N, 32, 32, (0, 1))
stream = torch.cuda.current_stream()
xsmm.set_stream(stream.cuda_stream)
graphs = []
for a, b in zip(A, B):
    with xsmm.graph_capture() as g:
        rp.run(a, b)
    graphs.append(g)
print("kernel:", xsmm.last_kernel())
for g in graphs:
    g.launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record(stream)
for _ in range(reps):
    for g in graphs:
        g.launch()
e1.record(stream)
torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3 / (reps * sets)
print(f"{t * 1e6:.2f} us per pack, {2 * M * N * 2 / t / 1e9:.0f} GB/s")
