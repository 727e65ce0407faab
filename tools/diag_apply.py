import os, sys, subprocess
code = r'''
import os, sys, torch
sys.path.insert(0, "/root/repo")
from tests.helpers import Config, rel_err
from tests.parity_cases import make_engine
name, B, w = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
cfg = Config(name)
x, adj, r2, flags = cfg.random_state(B, 1)
eng = make_engine(cfg, B, "cuda")
print(name, B, "net", w, eng.info(), flush=True)
ref = cfg.oracle_models[w](x, adj, r2, flags)
out = eng.score(w, x, adj, r2, flags).cpu()
print("  rel err", rel_err(out, ref), flush=True)
'''
open("/tmp/one.py", "w").write(code)
for args in [("enzymes_small_cc", 8, 0), ("enzymes_small_cc", 8, 1), ("enzymes_small_cc", 8, 2), ("enzymes_small_cc", 1, 2), ("community_small_cc", 2, 2), ("community_small_cc", 150, 2)]:
    r = subprocess.run([sys.executable, "/tmp/one.py"] + [str(a) for a in args], capture_output=True, text=True, timeout=120)
    print(r.stdout.strip()); 
    if r.returncode: print("  FAILED:", r.stderr.strip().splitlines()[-1][:200])
