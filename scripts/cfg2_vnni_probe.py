"""cfg2 (BRGEMM bf16 1024^3 x batch 16) with VNNI-2 packed B ([16][512][1024][2], the compiler's default bf16 layout):
direct invoke vs captured graph, against flat B. Timing + parity on a row slab."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tpp_mlir_b200 import xsmm

BF16 = xsmm.BF16
m = n = k = 1024
batch = 16
ns = 4
stream = torch.cuda.current_stream()
xsmm.set_stream(stream.cuda_stream)


def rnd(*s):
    return (torch.rand(*s, device="cuda") * 0.5).to(torch.bfloat16)


def timed(fns, iters):
    for f in fns:
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(iters):
        fns[i % len(fns)]()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


A = [rnd(batch, m, k) for _ in range(ns)]
B = [rnd(batch, k, n) for _ in range(ns)]
Bv = [b.reshape(batch, k // 2, 2, n).permute(0, 1, 3, 2).contiguous() for b in B]   # [b][k/2][n][2]
C = [torch.empty(m, n, dtype=torch.bfloat16, device="cuda") for _ in range(ns)]
flops = 2.0 * m * n * k * batch
for label, flags, Bs in (("flat", 4 | 64 | 128, B), ("vnni2", 4 | 2048 | 64 | 128, Bv)):
    h = xsmm.brgemm_dispatch(BF16, m, n, k, k, n, n, m * k, k * n, flags)
    fns = [lambda a=a, b=b, c=c: xsmm.LIB.xsmm_brgemm_invoke(BF16, h, a.data_ptr(), 0, b.data_ptr(), 0, c.data_ptr(), 0, batch)
           for a, b, c in zip(A, Bs, C)]
    t = timed(fns, 40)
    name = xsmm.last_kernel()
    want = torch.einsum("bik,bkj->ij", A[0][:, :8].double(), B[0].double())
    fns[0]()
    torch.cuda.synchronize()
    rel = float(((C[0][:8].double() - want).abs().max() / want.abs().max()).item())
    print(f"{label} direct: {t * 1e6:.1f} us, {flops / t / 1e12:.0f} TF/s, kernel {name}, rel err {rel:.2e}")
    graphs = []
    for a, b, c in zip(A, Bs, C):
        with xsmm.graph_capture() as g:
            xsmm.LIB.xsmm_brgemm_invoke(BF16, h, a.data_ptr(), 0, b.data_ptr(), 0, c.data_ptr(), 0, batch)
        graphs.append(g)
    name = xsmm.last_kernel()
    C[0].zero_()
    t = timed([g.launch for g in graphs], 40)
    graphs[0].launch()
    torch.cuda.synchronize()
    rel = float(((C[0][:8].double() - want).abs().max() / want.abs().max()).item())
    print(f"{label} captured: {t * 1e6:.1f} us, {flops / t / 1e12:.0f} TF/s, kernel {name}, rel err {rel:.2e}")
