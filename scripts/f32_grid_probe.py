"""fp32 MLP 3 x 1024^2, batch 256, --tiles=32,32,32 (the reference's benchmarks/config/base/base.json fp32 configs): 768
tile invokes per forward under capture. Time per forward and parity (f64-accumulate oracle, 1e-5)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import oracle
from tpp_mlir_b200 import harness, xsmm

F32 = 1
batch, layers, tiles = 256, (1024, 1024, 1024, 1024), (32, 32, 32)
bn, bk, bc = tiles
cfg = harness.MlpConfig(batch=batch, layers=layers, tiles=tiles, dtype=F32)
gen = oracle.TensorInit("normal", F32, 123)
Ws = [gen.fill(c, k) for c, k in zip(layers[:-1], layers[1:])]
bs = [gen.fill(k) for k in layers[1:]]
x = gen.fill(batch, layers[0])
t = torch.from_numpy
wp = [harness.pack_weight(t(W), bk, bc).cuda() for W in Ws]
acts = [harness.pack_activation(t(x), bn, bc).cuda()] + [torch.zeros(batch * k).cuda() for k in layers[1:]]
r = harness.MlpReplay(cfg, wp, [t(b).cuda() for b in bs], acts)
with xsmm.graph_capture() as g:
    r.forward()
print("kernel:", xsmm.last_kernel())
for _ in range(3):
    g.launch()
xsmm.sync()
n0 = xsmm.launch_count()
t0 = xsmm.perf_start_timer()
reps = 20
for _ in range(reps):
    g.launch()
dt = xsmm.perf_stop_timer(t0) / reps
print(f"{dt * 1e6:.1f} us per forward, {(xsmm.launch_count() - n0) / reps:.0f} launches per forward, {cfg.flops() / dt / 1e12:.2f} TF/s")
oracle.set_acc_mode(1)
ref = x
for W, b in zip(Ws, bs):
    y = np.zeros((batch, 1024), np.float32)
    oracle.fused_brgemm(F32, batch, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1, ref, W, y, b, 1)
    ref = y
got = harness.unpack_activation(acts[-1].reshape(batch // bn, 1024 // bk, bn, bk)).cpu().numpy()
print("max rel err", float(np.abs(got - ref).max() / np.abs(ref).max()))
