"""Timeline of the tensor-core apply kernel's pipeline roles (CTA 0): python tools/apply_trace.py"""
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
samp = sys.argv[3] if len(sys.argv) > 3 else "PC"
eng = make_engine(cfg, B, "cuda", sampler=samp)
eng.init(flags.cuda(), seed=1)
eng.run(0, 2)
tr = torch.zeros(512, 16, dtype=torch.int64, device="cuda")
nat.check(eng.lib.ccsd_debug_apply_trace(eng.handle, tr.data_ptr()))
eng.run(2, 3)   # the LAST apply pass of the step (PRED) leaves its stamps
torch.cuda.synchronize()
t = tr.cpu().numpy()
t0 = t[0, 0]
names = ["ld.issue", "ld.landed", "ld.opfree", "ld.full", "mma.start", "mma.issued", "e0.wait", "e0.acc", "e0.ld", "e0.done"]
print("tile " + " ".join(f"{n:>10s}" for n in names))
for gidx in list(range(0, 12)) + list(range(30, 44)) + list(range(100, 108)):
    print(f"{gidx:4d} " + " ".join(f"{(t[gidx, i] - t0):10d}" for i in range(10)))
d = np.diff(t[:250, 9])
print("epilogue(e0) tile period: mean", d.mean(), "median", np.median(d))
for a_, b_, nm in [(0, 1, "cp.async wait"), (1, 2, "opfree wait"), (2, 3, "convert"), (4, 5, "mma issue"), (6, 7, "e0 wait acc"), (7, 8, "e0 tmem ld"), (8, 9, "e0 compute")]:
    x = (t[2:250, b_] - t[2:250, a_])
    print(f"{nm:14s} mean {x.mean():9.0f} median {np.median(x):9.0f}")
x = t[2:250, 0][1:] - t[2:250, 3][:-1]
print(f"{'ld full->next issue (copy-out + cp.async issue)':14s} mean {x.mean():9.0f}")

nt = (cfg.K + 31) // 32
print("group boundaries (first tile of each group): e0.done(prev) -> fill.start -> fill.end -> mma.start -> e0.acc")
for gi in range(1, min(8, 500 // nt)):
    r = gi * nt
    print(f"tile {r:4d}: prev e0.done {t[r-1, 9] - t[r-1, 9]:7d}  fill.start {t[r, 14] - t[r-1, 9]:7d}  fill.end {t[r, 15] - t[r-1, 9]:7d}  "
          f"mma.start {t[r, 4] - t[r-1, 9]:7d}  e0.acc {t[r, 7] - t[r-1, 9]:7d}  next e0.done {t[r, 9] - t[r-1, 9]:7d}"
          f"  | refill done (warp 0) {t[r, 13] - t[r, 14]:6d}")
