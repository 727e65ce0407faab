"""Timeline of tc_gram_kernel (CTA 0): CCSD_B200_TRACE_GRAM=1 python tools/gram_trace.py [config] [B]"""
import os, sys, torch, numpy as np
os.environ["CCSD_B200_TRACE_GRAM"] = "1"
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
eng = make_engine(cfg, B, "cuda", predictor="Euler", corrector="None")   # one tc_gram and no traced tc_apply pass per step
eng.init(flags.cuda(), seed=1)
eng.run(0, 2)
tr = torch.zeros(512, 16, dtype=torch.int64, device="cuda")
nat.check(eng.lib.ccsd_debug_apply_trace(eng.handle, tr.data_ptr()))
eng.run(2, 3)
torch.cuda.synchronize()
t = tr.cpu().numpy()
nkb = (cfg.K + 63) // 64
t0 = t[0, 0]
print("kb   g0.start g0.issued g0.slot  g0.done | g1.start g1.issued g1.slot g1.done | mma.wait mma.full mma.issued | epi.tfull epi.m0 epi.m1   (cycles from the first stamp)")
for i in range(0, 3 * nkb + 4):
    r = t[i]
    f = lambda v: f"{v - t0:9d}" if v else "        -"
    print(f"{i:3d} " + " ".join(f(r[j]) for j in range(0, 4)) + " | " + " ".join(f(r[j]) for j in range(4, 8)) + " | " + " ".join(f(r[j]) for j in (8, 9, 10)) + " | " + " ".join(f(r[j]) for j in (12, 13, 14)))
ev = t[0:6 * nkb:2]
per = np.diff(ev[:, 0]); print("group 0 k-block period (2 k-blocks): median", np.median(per))
print("g0: issue->slot wait start", np.median(ev[:, 1] - ev[:, 0]), " slot wait", np.median(ev[:, 2] - ev[:, 1]), " convert+store+arrive", np.median(ev[:, 3] - ev[:, 2]))
mm = t[0:6 * nkb]
print("mma: wait for full", np.median(mm[:, 9] - mm[:, 8]), " issue", np.median(mm[:, 10] - mm[:, 9]), " period", np.median(np.diff(mm[:, 8])))

print("g0 section: slot->storeA", np.median(ev[:, 4] - ev[:, 2]), " loadA+storeB", np.median(ev[:, 5] - ev[:, 4]), " loadB", np.median(ev[:, 6] - ev[:, 5]),
      " P0 rows + ones", np.median(ev[:, 7] - ev[:, 6]), " fence+arrive", np.median(ev[:, 3] - ev[:, 7]))
