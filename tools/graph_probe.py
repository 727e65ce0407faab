"""How much would capturing a sampler step as a CUDA graph gain?  Captures ONE step (fixed step index: the replayed numbers
are not a valid trajectory, only the timing is meaningful) and compares K eager steps with K graph replays.
   python tools/graph_probe.py [config] [B]"""
import sys, time, torch
sys.path.insert(0, "/root/repo")
from tests.helpers import Config
from tests.parity_cases import make_engine
from ccsd_b200 import _native as nat
name = sys.argv[1] if len(sys.argv) > 1 else "community_small"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K = 200
cfg = Config(name)
g = torch.Generator().manual_seed(0)
n = torch.randint(max(2, cfg.N // 2), cfg.N + 1, (B,), generator=g)
flags = (torch.arange(cfg.N)[None, :] < n[:, None]).float()
eng = make_engine(cfg, B, "cuda")
eng.init(flags.cuda(), seed=1)
eng.run(0, 5)
torch.cuda.synchronize()
def timed(fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
t_eager = timed(lambda: eng.run(5, 5 + K))
s = torch.cuda.Stream()
gr = torch.cuda.CUDAGraph()
with torch.cuda.stream(s):
    eng.run(300, 301)
    torch.cuda.synchronize()
    gr.capture_begin()
    nat.check(eng.lib.ccsd_plan_run(eng.handle, 301, 302, torch.cuda.current_stream().cuda_stream))
    gr.capture_end()
torch.cuda.synchronize()
def replay():
    for _ in range(K): gr.replay()
t_graph = timed(replay)
print(f"{name} B={B}: eager {t_eager:.4f} ms/step, graph replay {t_graph:.4f} ms/step, ratio {t_eager / t_graph:.3f}")
