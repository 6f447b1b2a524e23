import sys, torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from tpp_mlir_b200 import xsmm
def run(m, n, k, b, iters=20):
    A = (torch.rand(b, m, k, device="cuda") - 0.5).bfloat16(); B = (torch.rand(b, k, n, device="cuda") - 0.5).bfloat16()
    C = torch.zeros(m, n, device="cuda", dtype=torch.bfloat16)
    h = xsmm.brgemm_dispatch(2, m, n, k, k, n, n, m * k, k * n, 4)
    st = torch.cuda.current_stream(); xsmm.set_stream(st.cuda_stream)
    for _ in range(3): xsmm.LIB.xsmm_brgemm_invoke(2, h, A.data_ptr(), 0, B.data_ptr(), 0, C.data_ptr(), 0, b)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(iters): xsmm.LIB.xsmm_brgemm_invoke(2, h, A.data_ptr(), 0, B.data_ptr(), 0, C.data_ptr(), 0, b)
    e1.record(st); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / iters
    print(f"{m}x{n}x{k} b{b}: {xsmm.last_kernel():45s} {2*m*n*k*b/t/1e12:8.1f} TF/s  {t*1e6:8.1f} us", flush=True)
if __name__ == "__main__":
  for shp in [(4096, 4096, 4096, 1), (8192, 8192, 2048, 1), (2048, 1024, 1024, 1), (1024, 1024, 1024, 16)]:
    run(*shp)
