"""One block-packed MLP forward under capture (tiles / vnni from argv), checked against the oracle; debugging aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch
import oracle
from tpp_mlir_b200 import harness, xsmm

tiles = tuple(int(v) for v in sys.argv[1].split(","))
vnni = sys.argv[2] == "1"
layers = tuple(int(v) for v in sys.argv[3].split(",")) if len(sys.argv) > 3 else (1024, 1024, 1024)
batch = 256
bn, bk, bc = tiles
cfg = harness.MlpConfig(batch=batch, layers=layers, tiles=tiles, vnni=vnni)
gen = oracle.TensorInit("normal", oracle.BF16, 123)
t = lambda a: torch.from_numpy(a.view(np.int16))
Ws = [gen.fill(c, k) for c, k in zip(layers[:-1], layers[1:])]
bs = [gen.fill(k) for k in layers[1:]]
x = gen.fill(batch, layers[0])
wp = [harness.pack_weight(t(W), bk, bc) for W in Ws]
if vnni:
    wp = [harness.vnni_pack_weight(w) for w in wp]
acts = [harness.pack_activation(t(x), bn, bc).cuda()] + [torch.zeros(batch * k, dtype=torch.int16).cuda() for k in layers[1:]]
r = harness.MlpReplay(cfg, [w.cuda() for w in wp], [t(b).cuda() for b in bs], acts)
with xsmm.graph_capture() as g:
    r.forward()
print("kernel:", xsmm.last_kernel(), flush=True)
g.launch()
xsmm.sync()
ref = x
for W, b in zip(Ws, bs):
    y = np.zeros((batch, W.shape[1]), np.uint16)
    oracle.fused_brgemm(2, batch, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, 5, 4, 1, ref, W, y, b, 1)
    ref = y
got = harness.unpack_activation(acts[-1].reshape(batch // bn, layers[-1] // bk, bn, bk)).cpu().numpy().view(np.uint16)
g32, w32 = oracle.bf16_to_f32(got), oracle.bf16_to_f32(ref)
err = np.abs(g32 - w32)
print("tiles", tiles, "vnni", vnni, "max rel err", err.max() / np.abs(w32).max(), "bad elems", int((err > 1e-2 * np.abs(w32).max()).sum()))
if err.max() > 1e-2 * np.abs(w32).max():
    bad = np.argwhere(err > 1e-2 * np.abs(w32).max())
    print("first bad", bad[:8].tolist(), "rows bad", np.unique(bad[:, 0])[:16], "cols bad", np.unique(bad[:, 1])[:32])
# replay timing (same buffers, L2-hot: what tpp-run's loop measures) and, with TPP_XSMM_TC_TRACE=4, the pair kernel's stamps
reps = int(os.environ.get("REPS", "200"))
for _ in range(10):
    g.launch()
xsmm.sync()
t0 = xsmm.perf_start_timer()
for _ in range(reps):
    g.launch()
dt = xsmm.perf_stop_timer(t0) / reps
print(f"replay: {dt * 1e6:.2f} us per forward ({cfg.flops() / dt / 1e12:.1f} TF/s)")
if os.environ.get("TPP_XSMM_TC_TRACE"):
    xsmm.LIB.xsmm_cuda_debug_dump_trace()
# UNROLL=k: tpp-run's loop unrolled k times before capture (exact repeats -> one sequential pass list)
unroll = int(os.environ.get("UNROLL", "0"))
if unroll > 1:
    stream = torch.cuda.current_stream()
    xsmm.set_stream(stream.cuda_stream)
    loop = harness.NativeMlpLoop(cfg, r.handles, [(r.acts, r.weights, r.biases)])
    loop.run_graph_unrolled(unroll * 4, unroll)
    xsmm.sync()
    n = unroll * 32
    t0 = xsmm.perf_start_timer()
    loop.run_graph_unrolled(n, unroll)
    dt = xsmm.perf_stop_timer(t0) / n
    print(f"unrolled x{unroll}: kernel {xsmm.last_kernel()}: {dt * 1e6:.2f} us per forward")
    if os.environ.get("TPP_XSMM_TC_TRACE"):
        xsmm.LIB.xsmm_cuda_debug_dump_trace()
