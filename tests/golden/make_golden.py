"""Generate the committed golden fixtures from the UNMODIFIED reference in /root/reference.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Writes into tests/golden/:
  weights_<cfg>.npz   the shipped checkpoint's weights + hyper-parameters + SDE/data config
  io_<cfg>.npz        reference outputs on seeded inputs: each score network, and a short sampler
                      run (the config's shipped sampler and the S4 / PC alternative) driven by
                      torch's CPU generator with a recorded seed.  Large rank-2 tensors are stored
                      as a strided sample plus sums.
  kat_reference_tests.npz
                      the reference's own known-answer tests (tests/models/test_ScoreNetwork_A_CC.py,
                      test_ScoreNetwork_F.py, test_hodge_layers.py, test_hodge_attention.py) replayed:
                      the module the test built (weights), the inputs the fixture produced, the full
                      reference output, and every (actual, expected, atol) triple the test asserted.
Nothing here is reference source: weights, inputs, outputs, and the expected numbers are data.
"""
from __future__ import annotations

import importlib.util
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import refstubs  # noqa: E402

refstubs.install()
from ccsd.src import solver as rsolver  # noqa: E402
from ccsd.src.utils import loader as rloader  # noqa: E402

OUT = Path(__file__).resolve().parent
STRIDE = 97  # sampling stride for big rank-2 tensors

CONFIGS = {
    # name: (checkpoint, shipped sampler (predictor, corrector, snr, scale_eps), B for io)
    "community_small": ("community_small/gdss_community_small", ("Euler", "Langevin", 0.05, 0.7), 3),
    "qm9": ("QM9/gdss_qm9_retrained", ("Reverse", "Langevin", 0.2, 0.7), 4),
    "qm9_cc": ("QM9/ccsd_qm9_CC", ("Reverse", "Langevin", 0.2, 0.7), 3),
    "community_small_cc": ("community_small_CC/ccsd_community_small_CC", ("Euler", "Langevin", 0.05, 0.7), 2),
    "enzymes_small_cc": ("ENZYMES_small_CC/ccsd_enzymes_small_CC", ("S4", "None", 0.15, 0.7), 2),
    # config/sample_ego_small_CC.yaml: Euler predictor, no corrector (snr / scale_eps unused)
    "ego_small_cc": ("ego_small_CC/ccsd_ego_small_CC", ("Euler", "None", 0.0, 0.0), 2),
    "ego_small": ("ego_small/gdss_ego_small", ("Euler", "None", 0.0, 0.0), 3),
    # ScoreNetworkA_Base_CC checkpoints (config/sample_*_Base_CC.yaml)
    "qm9_base_cc": ("QM9/ccsd_qm9_Base_CC", ("Reverse", "Langevin", 0.2, 0.7), 3),
    "community_small_base_cc": ("community_small_CC/ccsd_community_small_Base_CC", ("Euler", "Langevin", 0.05, 0.7), 2),
    "enzymes_small_base_cc": ("ENZYMES_small_CC/ccsd_enzymes_small_Base_CC", ("S4", "None", 0.15, 0.7), 2),
    "ego_small_cc_v2": ("ego_small_CC/ccsd_ego_small_CC_v2", ("Euler", "None", 0.0, 0.0), 2),
    # graph-only checkpoints of the remaining datasets
    "zinc250k": ("ZINC250k/gdss_zinc250k", ("Reverse", "Langevin", 0.2, 0.9), 3),
    "enzymes_small": ("ENZYMES_small/gdss_enzymes_small_retrained", ("S4", "None", 0.15, 0.7), 3),
    # large graphs (N > 64): no sample_*.yaml is shipped for them; GDSS's sampler settings
    "enzymes": ("ENZYMES/gdss_enzymes", ("Reverse", "Langevin", 0.1, 0.7), 2),
    "grid": ("grid/gdss_grid", ("Reverse", "Langevin", 0.1, 0.7), 1),
    # config/sample_grid_small_CC.yaml (N = 49, E = 1176, K = 18424: 87 MB of rank-2 state per sample)
    "grid_small_cc": ("grid_small_CC/ccsd_grid_small_CC", ("Reverse", "Langevin", 0.1, 0.7), 1),
}


def strip(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


def plain(d):
    return {k: (plain(v) if isinstance(v, dict) else v) for k, v in dict(d).items()}


def compress(t: torch.Tensor):
    """Full tensor when small, else strided sample + sums."""
    t = t.detach().to(torch.float32)
    if t.numel() <= 20000:
        return {"full": t.numpy()}
    f = t.reshape(-1)
    return {"sample": f[::STRIDE].numpy().copy(), "sum": np.float64(f.double().sum().item()),
            "abssum": np.float64(f.double().abs().sum().item()), "shape": np.array(t.shape)}


def put(store, key, t):
    for k, v in compress(t).items():
        store[f"{key}/{k}"] = v


# Model / layer options that no shipped checkpoint uses: ScoreNetworkX_GMH and the conv == "MLP"
# attention variant.  The reference's own classes with their default initialisation under torch.manual_seed(42).
SYNTH = {
    "synth_gmh_mlpconv": {
        "data": {"data": "synth", "max_node_num": 7, "max_feat_num": 3, "batch_size": 4},
        "sde": {k: {"type": "VP", "beta_min": 0.1, "beta_max": 1.0, "num_scales": 1000} for k in ("x", "adj")},
        "x": {"model_type": "ScoreNetworkX_GMH", "max_feat_num": 3, "depth": 2, "nhid": 8, "num_linears": 2, "c_init": 2,
              "c_hid": 4, "c_final": 3, "adim": 8, "num_heads": 4, "conv": "GCN"},
        "adj": {"model_type": "ScoreNetworkA", "max_feat_num": 3, "max_node_num": 7, "nhid": 8, "num_layers": 3,
                "num_linears": 2, "c_init": 2, "c_hid": 4, "c_final": 3, "adim": 8, "num_heads": 4, "conv": "MLP"},
        "sampler": ("Reverse", "Langevin", 0.1, 0.7), "B": 4,
    },
    "synth_gmh_mlpconv2": {   # GMH with the MLP convolution, three linears in the edge MLP
        "data": {"data": "synth", "max_node_num": 9, "max_feat_num": 4, "batch_size": 4},
        "sde": {k: {"type": "VE", "beta_min": 0.1, "beta_max": 1.0, "num_scales": 1000} for k in ("x", "adj")},
        "x": {"model_type": "ScoreNetworkX_GMH", "max_feat_num": 4, "depth": 3, "nhid": 12, "num_linears": 3, "c_init": 2,
              "c_hid": 5, "c_final": 4, "adim": 12, "num_heads": 4, "conv": "MLP"},
        "adj": {"model_type": "ScoreNetworkA", "max_feat_num": 4, "max_node_num": 9, "nhid": 12, "num_layers": 2,
                "num_linears": 3, "c_init": 2, "c_hid": 5, "c_final": 4, "adim": 12, "num_heads": 4, "conv": "GCN"},
        "sampler": ("Euler", "Langevin", 0.1, 0.7), "B": 4,
    },
}


def _synth_ckpt(name):
    from easydict import EasyDict
    sp = SYNTH[name]
    ck = {"model_config": EasyDict({"data": sp["data"], "sde": sp["sde"]})}
    torch.manual_seed(42)
    for k in ("x", "adj"):
        ck[f"params_{k}"] = dict(sp[k])
        m = rloader.load_model(dict(sp[k]))
        ck[f"{k}_state_dict"] = {kk: v.detach().clone() for kk, v in m.state_dict().items()}
    return ck


def make_config(name):
    if name in SYNTH:
        ckpt, (pred, corr, snr, seps), B = f"synthetic:{name} (torch.manual_seed(42) default init)", SYNTH[name]["sampler"], SYNTH[name]["B"]
        ck = _synth_ckpt(name)
    else:
        ckpt, (pred, corr, snr, seps), B = CONFIGS[name]
        ck = torch.load(f"/root/reference/checkpoints/{ckpt}.pth", map_location="cpu", weights_only=False)
    is_cc = "params_rank2" in ck
    keys = ["x", "adj"] + (["rank2"] if is_cc else [])
    cfg = ck["model_config"]
    meta = {"is_cc": is_cc, "data": plain(cfg.data), "sde": plain(cfg.sde), "params": {}, "checkpoint": ckpt,
            "shipped_sampler": {"predictor": pred, "corrector": corr, "snr": snr, "scale_eps": seps}}
    w = {}
    models = []
    for k in keys:
        p = plain(ck[f"params_{k}"])
        meta["params"][k] = p
        sd = strip(ck[f"{k}_state_dict"])
        for kk, v in sd.items():
            w[f"{k}/{kk}"] = v.detach().to(torch.float32).numpy()
        m = rloader.load_model(dict(p))
        m.load_state_dict(sd)
        m.eval()
        models.append(m)
    w["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(OUT / f"weights_{name}.npz", **w)

    # ---- reference I/O ----
    d = cfg.data
    N, Fd = d.max_node_num, d.max_feat_num
    d_min, d_max = (d.d_min, d.d_max) if is_cc else (None, None)
    g = torch.Generator().manual_seed(1234)
    n = torch.randint(max(2, N // 2), N + 1, (B,), generator=g)
    n[0] = N  # one full graph
    flags = (torch.arange(N)[None, :] < n[:, None]).to(torch.float32)
    x = torch.randn(B, N, Fd, generator=g) * flags[:, :, None]
    a = torch.randn(B, N, N, generator=g).triu(1)
    adj = (a + a.transpose(-1, -2)) * flags[:, :, None] * flags[:, None, :]
    io = {"flags": flags.numpy(), "x": x.numpy(), "adj": adj.numpy(), "seed_inputs": np.int64(1234)}
    args = [x, adj]
    if is_cc:
        from ccsd.src.utils.cc_utils import get_rank2_dim, mask_rank2
        E, K = get_rank2_dim(N, d_min, d_max)
        r2 = mask_rank2(torch.randn(B, E, K, generator=g) * 0.3, N, d_min, d_max, flags)
        io["rank2_seeded_scale"] = np.float32(0.3)
        args.append(r2)
    args.append(flags)
    with torch.no_grad():
        for k, m in zip(keys, models):
            put(io, f"net_{k}", m(*args))
    sdes = [rloader.load_sde(cfg.sde[k]) for k in keys]
    shapes = [(B, N, Fd), (B, N, N)] + ([(B, E, K)] if is_cc else [])
    runs = [(pred, corr)] + ([("S4", "None")] if pred != "S4" else [("Reverse", "Langevin")])
    steps = 3
    import ccsd.src.solver as S
    for (p_, c_) in runs:
        kw = dict(predictor=p_, corrector=c_, snr=snr, scale_eps=seps, n_steps=1, probability_flow=False,
                  continuous=True, denoise=True, eps=1e-4, device="cpu")
        if is_cc:
            kw.update(is_cc=True, sde_rank2=sdes[2], shape_rank2=shapes[2], d_min=d_min, d_max=d_max)
        fac = rsolver.S4_solver if p_ == "S4" else rsolver.get_pc_sampler
        orig = S.trange
        S.trange = lambda a_, b_, **k_: range(a_, min(b_, steps))
        try:
            fn = fac(sdes[0], sdes[1], shapes[0], shapes[1], **kw)
            torch.manual_seed(4321)
            out = fn(*models, flags)
        finally:
            S.trange = orig
        tag = f"run_{p_}_{c_}"
        io[f"{tag}/seed"] = np.int64(4321)
        io[f"{tag}/steps"] = np.int64(steps)
        for k, t in zip(keys, out[: len(keys)]):
            put(io, f"{tag}/{k}", t)
    np.savez_compressed(OUT / f"io_{name}.npz", **io)
    print("wrote", name, {k: int(sum(v.size for kk, v in w.items() if kk.startswith(k + "/"))) for k in keys})


# ---- the reference's own known-answer tests, replayed with recording ----
KAT_TESTS = {
    "test_ScoreNetwork_A_CC.py": ["test_ScoreNetworkA_CC"],
    "test_ScoreNetwork_F.py": ["test_ScoreNetworkF"],
    "test_hodge_layers.py": ["test_DenseHCNConv", "test_HodgeNetworkLayer", "test_BaselineBlock", "test_HodgeBaselineLayer"],
    "test_hodge_attention.py": ["test_HodgeAttention", "test_HodgeAdjAttentionLayer"],
    "test_ScoreNetwork_A_Base_CC.py": ["test_ScoreNetworkA_Base_CC"],
}
KAT_CLASSES = {"ScoreNetworkA_CC", "ScoreNetworkF", "DenseHCNConv", "HodgeNetworkLayer", "HodgeAttention",
               "HodgeAdjAttentionLayer", "BaselineBlock", "HodgeBaselineLayer", "ScoreNetworkA_Base_CC"}


def _raw(fixture):
    for attr in ("_get_wrapped_function", ):
        if hasattr(fixture, attr):
            return getattr(fixture, attr)()
    return getattr(fixture, "__wrapped__", fixture)


def make_kats():
    import inspect
    store = {}
    index = []
    for fname, tests in KAT_TESTS.items():
        spec = importlib.util.spec_from_file_location("reftest_" + fname[:-3], f"/root/reference/tests/models/{fname}")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)  # seeds torch / numpy with 42 like a fresh pytest import
        for tname in tests:
            fn = getattr(mod, tname)
            argn = list(inspect.signature(fn).parameters)
            np.random.seed(42)  # state the fixture sees when the file is run on its own
            cache = {}

            def resolve(name):  # function-scoped fixtures, dependencies first, each built once
                if name not in cache:
                    raw = _raw(getattr(mod, name))
                    deps = list(inspect.signature(raw).parameters)
                    cache[name] = raw(**{dn: resolve(dn) for dn in deps})
                return cache[name]

            fixtures = {a: resolve(a) for a in argn}
            calls, asserts = [], []
            depth = [0]
            orig_call = torch.nn.Module.__call__
            orig_allclose = torch.allclose

            def call(self, *a, **k):
                depth[0] += 1
                try:
                    out = orig_call(self, *a, **k)
                finally:
                    depth[0] -= 1
                if depth[0] == 0 and type(self).__name__ in KAT_CLASSES:
                    calls.append((self, a, k, out))
                return out

            def allclose(a_, b_, **k):
                asserts.append((a_.detach().clone(), b_.detach().clone(), k.get("atol", 1e-8)))
                return orig_allclose(a_, b_, **k)

            torch.nn.Module.__call__ = call
            torch.allclose = allclose
            try:
                fn(**fixtures)  # raises if the reference fails its own KAT here
            finally:
                torch.nn.Module.__call__ = orig_call
                torch.allclose = orig_allclose
            # the KAT assertions concern the LAST top-level call of the test
            m, a, k, out = calls[-1]
            entry = {"test": tname, "file": fname, "cls": type(m).__name__,
                     "hp": {kk: vv for kk, vv in vars(m).items() if isinstance(vv, (int, float, str, bool))},
                     "n_args": len(a), "kwargs": {kk: (None if vv is None else "tensor") for kk, vv in k.items()},
                     "n_asserts": len(asserts)}
            for kk, vv in m.state_dict().items():
                store[f"{tname}/sd/{kk}"] = vv.detach().numpy()
            for i, t in enumerate(a):
                if torch.is_tensor(t):
                    store[f"{tname}/arg{i}"] = t.detach().numpy()
                else:
                    entry[f"arg{i}"] = t
            outs = out if isinstance(out, (tuple, list)) else [out]
            for i, t in enumerate(outs):
                store[f"{tname}/out{i}"] = t.detach().numpy()
            for i, (act, exp, atol) in enumerate(asserts):
                store[f"{tname}/assert{i}/actual"] = act.numpy()
                store[f"{tname}/assert{i}/expected"] = exp.numpy()
                store[f"{tname}/assert{i}/atol"] = np.float64(atol)
            index.append(entry)
            print("KAT", tname, "calls", len(calls), "asserts", len(asserts))
    store["index"] = np.frombuffer(json.dumps(index).encode(), dtype=np.uint8)
    np.savez_compressed(OUT / "kat_reference_tests.npz", **store)


if __name__ == "__main__":
    which = sys.argv[1:] or (list(CONFIGS) + ["kats"])
    for n_ in which:
        if n_ == "kats":
            make_kats()
        else:
            make_config(n_)
