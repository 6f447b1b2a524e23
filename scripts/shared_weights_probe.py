"""The served-model form of the headline workload (bench.py: extra.shared_weights_batch256): 148 forward passes of batch
256 per launch on ONE set of weights; a profiling target for ncu (tensor-pipe share of the pair-per-chain kernel when
the weights are L2-resident)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from tpp_mlir_b200 import xsmm

gen, Ws, bs = bench.make_host_data()
x = gen.fill(256, 1024)
dev = torch.device("cuda", 0)
to_dev = lambda a: torch.from_numpy(a.view(np.int16)).to(dev)
stream = torch.cuda.current_stream()
xsmm.set_stream(stream.cuda_stream)
# argv: [tiles "bn,bk,bc"] [vnni 0|1], e.g. "32,32,32 1" = the reference's default stream as a served model
tiles = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else None
vnni = len(sys.argv) > 2 and sys.argv[2] == "1"
wl = bench.MlpWorkload(256, to_dev(x), [to_dev(W) for W in Ws], [to_dev(b) for b in bs], shared_weights=True, min_sets=126,
                       tiles=tiles, vnni=vnni)
wl.rotations(3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
wl.rotations(10)
e1.record(stream)
torch.cuda.synchronize()
t = e0.elapsed_time(e1) * 1e-3 / (10 * wl.num_sets)
print(f"{xsmm.last_kernel()}: {t * 1e6:.3f} us per forward, {wl.cfg.flops() / t / 1e12:.0f} TF/s, {wl.num_sets} operand sets")
