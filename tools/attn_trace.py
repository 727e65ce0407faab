"""Phase timeline of tc_attn_kernel (CTA 0, layer 1 of the predictor's evaluation): python tools/attn_trace.py [config] [B]"""
import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from tests.helpers import Config
from tests.parity_cases import make_engine
from ccsd_b200 import _native as nat
name = sys.argv[1] if len(sys.argv) > 1 else "community_small_cc"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = Config(name)
g = torch.Generator().manual_seed(0)
n = torch.randint(max(2, cfg.N // 2), cfg.N + 1, (B,), generator=g)
flags = (torch.arange(cfg.N)[None, :] < n[:, None]).float()
eng = make_engine(cfg, B, "cuda")
eng.init(flags.cuda(), seed=1)
eng.run(0, 2)
tr = torch.zeros(512, 16, dtype=torch.int64, device="cuda")
nat.check(eng.lib.ccsd_debug_apply_trace(eng.handle, tr.data_ptr()))
eng.run(2, 3)
torch.cuda.synchronize()
t = tr.cpu().numpy()
names = ["item", "La+Lb", "Lc", "Ld", "MMA-A", "E1", "MMA-B", "E2", "S", "sym"]
print("setup (kernel start -> first item):", t[0, 0] - t[0, 15], "cycles")
print("item " + " ".join(f"{n:>8s}" for n in names[1:]) + "    total")
for it in range(32):
    if t[it, 9] == 0 or t[it, 9] < t[it, 0] or t[it, 9] - t[it, 0] > 10 ** 6: break
    d = [t[it, i + 1] - t[it, i] for i in range(9)]
    print(f"{it:4d} " + " ".join(f"{x:8d}" for x in d) + f" {t[it, 9] - t[it, 0]:8d}")
print("kernel (start -> last stamp):", t[:, 9].max() - t[0, 15], "cycles")
