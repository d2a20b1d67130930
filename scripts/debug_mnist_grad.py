import faulthandler; faulthandler.dump_traceback_later(60, exit=True)
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_gpu_mnist as t
conf = {"alg_name": "dsgd", "alpha0": 0.01, "mu": 0.001, "outer_iterations": 2, "profile": False}
fused = t._problem(3, 64, "fused", conf, M=150)
ref = t._problem(3, 64, "torch", conf, M=150)
ref.arena.theta.copy_(fused.arena.theta)
for step in range(3):
    fused.compute_grads(); ref.compute_grads()
    d = (fused.arena.grad - ref.arena.grad).abs()
    bad = d > 2e-5 + 2e-3 * ref.arena.grad.abs()
    print("step", step, "bad", int(bad.sum()), "maxabs", d.max().item())
    for s in fused.layout.slots:
        sl = slice(s.offset, s.offset + s.numel)
        b = bad[:, sl]
        if b.any():
            idx = b.nonzero()
            print("  slot", s.name, "bad", int(b.sum()), "first", idx[:5].tolist(), "last", idx[-3:].tolist())
            if s.name == "seq.4.weight":
                rows = (idx[:, 1] // 432).unique().tolist(); cols = (idx[:, 1] % 432)
                print("   rows", rows[:20], "cols min/max", cols.min().item(), cols.max().item(), "nodes", idx[:,0].unique().tolist())
                l, o = idx[0].tolist()
                print("   sample vals fused/ref", fused.arena.grad[l, s.offset+o].item(), ref.arena.grad[l, s.offset+o].item())
