"""Reference-side samples for the metric-level check (BASELINE.json north_star: "the repo's MMD / validity metrics within
run-to-run variance"): FULL 1000-step sampler runs of the oracle (oracle/ccsd_oracle.py, validated against the unmodified
reference by oracle/validate_against_reference.py) on the CPU with torch noise, several seeds, shipped checkpoints and
shipped sampler settings.  The quantised samples are committed bit-packed as tests/golden/metric_samples_<cfg>.npz; the
GPU test (tests/test_metrics_gpu.py) runs the CUDA sampler with Philox noise on the same flags and compares graph
statistics with the seed-to-seed spread of these runs.

    python tests/golden/make_golden_metrics.py community_small 64 0 1 2        (minutes of CPU per seed)
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ccsd_oracle as O  # noqa: E402
from tests.helpers import Config  # noqa: E402


def flags_for(N: int, B: int) -> torch.Tensor:
    rng = np.random.RandomState(0)
    n = rng.randint((N + 1) // 2, N + 1, size=B)
    return torch.from_numpy((np.arange(N)[None, :] < n[:, None]).astype(np.float32))


def main():
    name, B = sys.argv[1], int(sys.argv[2])
    seeds = [int(s) for s in sys.argv[3:]]
    cfg = Config(name)
    flags = flags_for(cfg.N, B)
    sh = cfg.shipped
    out = {"flags": flags.numpy().astype(np.uint8), "seeds": np.array(seeds), "B": B}
    for seed in seeds:
        t = time.time()
        kw = dict(snr=sh["snr"], scale_eps=sh["scale_eps"], denoise=True, eps=1e-4, d_min=cfg.d_min, d_max=cfg.d_max,
                  noise=O.NoiseSource(seed))
        if sh["predictor"] == "S4":
            res, _ = O.s4_solver(cfg.oracle_models, cfg.sdes(), cfg.shapes(B), flags, **kw)
        else:
            res, _ = O.pc_sampler(cfg.oracle_models, cfg.sdes(), cfg.shapes(B), flags, predictor=sh["predictor"],
                                  corrector=sh["corrector"], n_steps=1, **kw)
        mol = name.startswith("qm9") or name.startswith("zinc")
        out[f"x_{seed}"] = res[0].numpy().astype(np.float16)
        if mol:
            out[f"adj_{seed}"] = O.quantize_mol(res[1]).numpy().astype(np.uint8)
        else:
            out[f"adj_{seed}"] = np.packbits(O.quantize(res[1]).numpy().astype(np.uint8))
        if cfg.is_cc:
            out[f"rank2_{seed}"] = np.packbits(O.quantize(res[2]).numpy().astype(np.uint8))
        print(name, "seed", seed, f"{time.time() - t:.0f}s", flush=True)
    np.savez_compressed(Path(__file__).resolve().parent / f"metric_samples_{name}.npz", **out)


if __name__ == "__main__":
    main()
