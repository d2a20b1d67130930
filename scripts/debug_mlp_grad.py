import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_mlp as t
for h1, B, loss in [(64, 300, "MSE"), (64, 300, "BCE"), (256, 300, "MSE"), (128, 1500, "L1")]:
    fused = t._density_problem("fused", h1=h1, B=B, loss=loss); ref = t._density_problem("torch", h1=h1, B=B, loss=loss)
    ref.arena.theta.copy_(fused.arena.theta)
    lf = fused.compute_grads().clone(); lr = ref.compute_grads().clone()
    print(h1, B, loss, "loss", lf.tolist(), lr.tolist())
    for s in fused.layout.slots:
        a = fused.arena.grad[:, s.offset: s.offset + s.numel]; b = ref.arena.grad[:, s.offset: s.offset + s.numel]
        cos = torch.nn.functional.cosine_similarity(a.reshape(1, -1), b.reshape(1, -1)).item()
        print("   ", s.name, "rel", ((a - b).norm() / b.norm()).item(), "cos", cos, "norm", b.norm().item())
