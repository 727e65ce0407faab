"""GPU check of the tcgen05 attention-channel kernel (tc_attn.cuh): adjacency-score parity against the oracle for
graph sizes / batch sizes that exercise full and short groups, with the kernel on and (A/B) off."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.parity_cases import score_parity, make_engine
from tests.helpers import Config

cases = [("community_small", 16), ("community_small", 6), ("community_small", 1), ("qm9", 64), ("qm9", 15), ("community_small_cc", 4),
         ("qm9_cc", 16), ("enzymes_small_cc", 8), ("ego_small", 8), ("zinc250k", 8), ("enzymes_small", 16), ("ego_small_cc", 2),
         ("grid_small_cc", 1)]
only = sys.argv[1:] 
for name, B in cases:
    if only and name not in only:
        continue
    t = time.time()
    try:
        cfg = Config(name)
        eng = make_engine(cfg, B, "cuda")
        ntc = tuple(eng.lib.ccsd_plan_info(eng.handle, i) for i in (14, 15, 16, 17))
        errs = score_parity(name, B, "cuda")
        print(f"{name:24s} B={B:4d} tc(attn,xfin,hnorm,edge)={ntc} errs={ {k: float('%.3g' % v) for k, v in errs.items()} } {time.time()-t:.1f}s", flush=True)
    except Exception as e:
        print(f"{name:24s} B={B:4d} FAILED: {type(e).__name__}: {e}", flush=True)
        if "CUDA" in str(e) or "cuda" in str(e):
            break
