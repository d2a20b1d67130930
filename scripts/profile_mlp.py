"""One fused tcgen05 MLP training step on the online-density shape, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_gpu_mlp as t
pr = t._density_problem("fused", h1=256, B=12500, M=200000, N=7)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    pr.fused.launch()
torch.cuda.synchronize()
print("done")
