"""Print score / sampler-step parity errors for a config (GPU):  python tools/diag_parity.py qm9_base_cc PC Reverse Langevin 8"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tests.parity_cases import sampler_parity, score_parity

name, sampler, pred, corr, B = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5])
print(name, "scores", {k: f"{v:.2e}" for k, v in score_parity(name, B, "cuda").items()})
for steps in (1, 2, 3, 5):
    res = sampler_parity(name, sampler, pred, corr, B, steps, "cuda")
    print(name, "steps", steps, {k: tuple(f"{e:.2e}" for e in v) for k, v in res.items()})
