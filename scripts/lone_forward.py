"""Lone forward pass (one chain, same buffers, graph replay): per-forward time and, with TPP_XSMM_TC_TRACE=3, the pass
kernel's clock-stamp trace. Debugging / tuning aid for the `latency` field of bench.py."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from tpp_mlir_b200 import harness, xsmm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
layers = (1024, 1024, 1024, 1024)
cfg = harness.MlpConfig(batch=256, layers=layers, tiles=(256, 1024, 1024))
gen = oracle.TensorInit("normal", oracle.BF16, 123)
t = lambda a: torch.from_numpy(a.view(np.int16)).cuda()
Ws = [gen.fill(1024, 1024) for _ in range(3)]
bs = [gen.fill(1024) for _ in range(3)]
x = gen.fill(256, 1024)
acts = [t(x)] + [torch.zeros(256 * 1024, dtype=torch.int16).cuda() for _ in range(3)]
r = harness.MlpReplay(cfg, [t(W) for W in Ws], [t(b) for b in bs], acts)
stream = torch.cuda.current_stream()
xsmm.set_stream(stream.cuda_stream)
loop = harness.NativeMlpLoop(cfg, r.handles, [(acts, r.weights, r.biases)])
unroll = int(os.environ.get("UNROLL", "1"))
run = (lambda k: loop.run_graph_unrolled(k, unroll)) if unroll > 1 else loop.run_graph
run(64)
torch.cuda.synchronize()
t0 = xsmm.perf_start_timer()
run(n)
dt = xsmm.perf_stop_timer(t0) / n
print(f"kernel {xsmm.last_kernel()}: {dt * 1e6:.2f} us per forward, {cfg.flops() / dt / 1e12:.1f} TF/s")
if os.environ.get("TPP_XSMM_TC_TRACE"):
    xsmm.LIB.xsmm_cuda_debug_dump_trace()
ref = x
for W, b in zip(Ws, bs):
    y = np.zeros((256, 1024), np.uint16)
    oracle.fused_brgemm(2, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1, ref, W, y, b, 1)
    ref = y
g32, w32 = oracle.bf16_to_f32(acts[-1].cpu().numpy().view(np.uint16).reshape(256, 1024)), oracle.bf16_to_f32(ref)
print("max rel err", np.abs(g32 - w32).max() / np.abs(w32).max())
