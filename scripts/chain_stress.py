import sys, numpy as np, torch
sys.path.insert(0, ".")
import oracle
from tpp_mlir_b200 import xsmm
gen = oracle.TensorInit("normal", 2, 5)
L = 3
dev = lambda a: torch.from_numpy(a.view(np.int16)).cuda()
Ws = [dev(gen.fill(1024, 1024)) for _ in range(L)]; bs = [dev(gen.fill(1024)) for _ in range(L)]
h = xsmm.fused_brgemm_dispatch(2, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1)
bad = 0
for trial in range(40):
    x = dev(gen.fill(256, 1024))
    acts = [x] + [torch.zeros(256, 1024, dtype=torch.int16, device="cuda") for _ in range(L)]
    def fwd():
        for l in range(L): xsmm.fused_brgemm_invoke(2, h, acts[l], 0, Ws[l], 0, acts[l + 1], 0, bs[l], 0, 1)
    fwd(); xsmm.sync(); want = acts[-1].clone()
    with xsmm.graph_capture() as g: fwd()
    for rep in range(50):
        acts[1].fill_(7); acts[2].fill_(9); acts[3].zero_()
        g.launch()
    xsmm.sync()
    if not torch.equal(acts[-1], want): bad += 1
    g.destroy()
print("chain stress: mismatches", bad, "of 40 trials x 50 replays")
