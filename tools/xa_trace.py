"""Per-phase cycle split of xa_kernel for graph 0 (clock64 stamps): python tools/xa_trace.py [config] [B]"""
import sys, torch
sys.path.insert(0, "/root/repo")
from tests.helpers import Config
from tests.parity_cases import make_engine
from ccsd_b200 import _native as nat
name = sys.argv[1] if len(sys.argv) > 1 else "community_small_cc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = Config(name)
g = torch.Generator().manual_seed(0)
n = torch.randint(max(2, cfg.N // 2), cfg.N + 1, (B,), generator=g); n[0] = cfg.N
flags = (torch.arange(cfg.N)[None, :] < n[:, None]).float()
pred, corr = cfg.shipped["predictor"], cfg.shipped["corrector"]
eng = make_engine(cfg, B, "cuda", sampler="S4" if pred == "S4" else "PC", predictor=pred if pred != "S4" else "Euler", corrector=corr)
eng.init(flags.cuda(), seed=1)
eng.run(0, 2)
tr = torch.zeros(8192 + 64, dtype=torch.int64, device="cuda")
nat.check(eng.lib.ccsd_debug_apply_trace(eng.handle, tr.data_ptr()))
eng.run(2, 3)
torch.cuda.synchronize()
t = tr.cpu().numpy()[8192:8192 + 16]
names = ["load", "X gcn layers", "X final MLP", "pow_tensor", "A layer 0", "A layer 1 channel loop", "A layer 1 node MLP", "A layer 1 edge MLP", "A layers 2..", "hodge branch", "A final MLP"]
tot = t[10] - t[0]
print(eng.info())
prev = t[0]
for i, nm in enumerate(names):
    if i == 0: continue
    d = t[i] - t[i - 1]
    print(f"{names[i] if i < len(names) else i:28s} {d:9d} cycles  {100 * d / tot:5.1f}%")
print("total", tot, "cycles =", tot / 1.965e3, "us for graph 0 (CTA 0, first wave)")
