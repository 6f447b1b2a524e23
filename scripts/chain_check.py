"""Chain-kernel check: parity vs the oracle, difference vs the per-layer kernels, replay determinism with
poisoned intermediates, and per-forward time of the graph-replayed chain (rotating operand sets > L2).
    python scripts/chain_check.py [layers] [trials]
TPP_XSMM_CHAIN=s selects the split-K chain kernel, =0 disables chain fusion (three PDL-chained kernels)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import oracle
from tpp_mlir_b200 import xsmm

L = int(sys.argv[1]) if len(sys.argv) > 1 else 3
TRIALS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
gen = oracle.TensorInit("normal", 2, 5)
dev = lambda a: torch.from_numpy(a.view(np.int16)).cuda()
hW = [gen.fill(1024, 1024) for _ in range(L)]
hb = [gen.fill(1024) for _ in range(L)]
Ws, bs = [dev(w) for w in hW], [dev(b) for b in hb]
h = xsmm.fused_brgemm_dispatch(2, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1)
stream = torch.cuda.current_stream()
xsmm.set_stream(stream.cuda_stream)
bad = 0
for trial in range(TRIALS):
    hx = gen.fill(256, 1024)
    x = dev(hx)
    acts = [x] + [torch.zeros(256, 1024, dtype=torch.int16, device="cuda") for _ in range(L)]

    def fwd():
        for l in range(L):
            xsmm.fused_brgemm_invoke(2, h, acts[l], 0, Ws[l], 0, acts[l + 1], 0, bs[l], 0, 1)

    fwd()
    xsmm.sync()
    direct = [a.clone() for a in acts[1:]]
    with xsmm.graph_capture() as g:
        fwd()
    name = xsmm.last_kernel()
    first = None
    for rep in range(50):
        for a, v in zip(acts[1:], (7, 9, 11, 13)):
            a.fill_(v)
        g.launch()
        if rep == 0:
            xsmm.sync()
            first = [a.clone() for a in acts[1:]]
    xsmm.sync()
    same = all(torch.equal(a, f) for a, f in zip(acts[1:], first))
    if not same:
        bad += 1
    if trial == 0:
        ref = hx
        for W, b in zip(hW, hb):
            y = np.zeros((256, 1024), np.uint16)
            oracle.fused_brgemm(2, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1, ref, W, y, b, 1)
            ref = y
        got = oracle.bf16_to_f32(acts[-1].cpu().numpy().view(np.uint16))
        want = oracle.bf16_to_f32(ref)
        rel = np.abs(got - want).max() / np.abs(want).max()
        dd = oracle.bf16_to_f32(direct[-1].cpu().numpy().view(np.uint16))
        reld = np.abs(got - dd).max() / np.abs(dd).max()
        nd = int((acts[-1] != direct[-1]).sum())
        print(f"kernel={name} rel_err_vs_oracle={rel:.3e} rel_diff_vs_per_layer={reld:.3e} differing_elems={nd}")
    g.destroy()
print(f"replay determinism: {bad} unstable of {TRIALS} trials x 50 replays")

# timing: 17 rotating operand sets (> L2), one graph holding all 17 forwards
sets = []
for s in range(17):
    a = [dev(gen.fill(256, 1024))] + [torch.zeros(256, 1024, dtype=torch.int16, device="cuda") for _ in range(L)]
    sets.append((a, [w.clone() for w in Ws], [b.clone() for b in bs]))
with xsmm.graph_capture() as g:
    for a, W, B in sets:
        for l in range(L):
            xsmm.fused_brgemm_invoke(2, h, a[l], 0, W[l], 0, a[l + 1], 0, B[l], 0, 1)
multi_name = xsmm.last_kernel()
for _ in range(5):
    g.launch()
xsmm.sync()
# every chain of the multi-chain launch against the same chain run alone (per-layer kernels): <= 1 bf16 ulp
worst, ndiff = 0, 0
for a, W, B in sets:
    got = [t.clone() for t in a[1:]]
    for l in range(L):
        xsmm.fused_brgemm_invoke(2, h, a[l], 0, W[l], 0, a[l + 1], 0, B[l], 0, 1)
    xsmm.sync()
    for t, u in zip(got, a[1:]):
        d = (t.cpu().numpy().view(np.uint16).astype(np.int32) - u.cpu().numpy().view(np.uint16).astype(np.int32))
        worst = max(worst, int(np.abs(d).max()))
        ndiff += int((d != 0).sum())
print(f"multi-chain launch {multi_name}: max |ulp diff| vs per-layer kernels over 17 chains x {L} layers = {worst}, "
      f"differing elements = {ndiff}")
for a, W, B in sets:
    for t in a[1:]:
        t.fill_(0x7FC0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
R = 60
for _ in range(R):
    g.launch()
e1.record(stream)
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (R * 17)
print(f"{xsmm.last_kernel()}: {us:.2f} us per {L}-layer forward "
      f"({(2*256*1024*1024+2*256*1024)*L/us/1e6:.1f} TFLOP/s), launches in graph: see launch_count")
import os
if os.environ.get("TPP_XSMM_TC_TRACE") == "3":
    xsmm.LIB.xsmm_cuda_debug_dump_trace()
