import os, sys, torch
sys.path.insert(0, "/root/repo")
from tests.helpers import Config, rel_err
from tests.test_tc_apply_gpu import _engine
for name, B in [("enzymes_small_cc", 9), ("qm9_cc", 16), ("community_small_cc", 2)]:
    cfg = Config(name)
    for scale in (0.5, 0.1):
        x, adj, r2, flags = cfg.random_state(B, seed=4, r2_scale=scale)
        ref = cfg.oracle_models[2](x, adj, r2, flags)
        a = _engine(cfg, B, True).score(2, x, adj, r2, flags).cpu()
        e = _engine(cfg, B, False)
        b = e.score(2, x, adj, r2, flags).cpu()
        H0, _ = e.debug_gram(r2, False); H1, _ = e.debug_gram(r2, True)
        hf = (H0.cpu().double() @ r2.double())
        print(name, scale, "fmode", "simt", rel_err(a, ref), "tc", rel_err(b, ref), "H tc-vs-simt", rel_err(H1, H0),
              "|ref|max", ref.abs().max().item(), "|hf|max", hf.abs().max().item(), "bad frac", ((b-ref).abs() > 1e-4*ref.abs().max()).float().mean().item())
