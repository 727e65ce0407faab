#!/usr/bin/env python
"""bench.py -- headline benchmark: sampled complexes/sec of the full PC sampler.

Contract (driver): ``python bench.py --gpus N --steps K --warmup W`` (torchrun for N > 1) prints ONE
JSON line.  Workload = BASELINE.json configs[1]: community_small_CC (N=20, F=11, E=190, K=1140),
ScoreNetworkX + ScoreNetworkA_CC + ScoreNetworkF, VP x3, Euler predictor + Langevin corrector
(snr .05, scale_eps .7), batch 1024 PER GPU, shipped-checkpoint weights (tests/golden), synthetic
node-count flags, Philox noise.  A "step" is one sampler iteration (corrector + predictor, 2 score
triples) over the whole batch on the real 1000-step schedule; ``value`` = complexes/sec of the full
1000-step sampler = B_total / (ms_per_step * 1000 steps).  With the default K = 1000 the timed
region IS the whole sampler run (prior sampling included).

``--impl reference`` times the reference algorithm's CPU port (oracle/ccsd_oracle.py, validated
against the unmodified reference in the build container -- the reference itself is Python and
cannot travel to the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (golden weights, per-GPU batch, CPU-baseline batch [reference's own per-call batch, SURVEY 8d])
    "community_small_cc": ("community_small_cc", 1024, 32),
    "qm9_cc": ("qm9_cc", 10000, 2500),
    "enzymes_small_cc": ("enzymes_small_cc", 4096, 64),
    "community_small": ("community_small", 128, 128),
    "ego_small": ("ego_small", 128, 128),
    "qm9_base_cc": ("qm9_base_cc", 10000, 2500),
    "community_small_base_cc": ("community_small_base_cc", 1024, 32),
    "ego_small_cc": ("ego_small_cc", 128, 32),
    "ego_small_cc_v2": ("ego_small_cc_v2", 128, 32),
    "enzymes_small_base_cc": ("enzymes_small_base_cc", 4096, 64),
    "zinc250k": ("zinc250k", 10000, 2500),
    "enzymes_small": ("enzymes_small", 64, 64),
    "enzymes": ("enzymes", 64, 8),
    "grid": ("grid", 64, 2),
    "grid_small_cc": ("grid_small_cc", 64, 1),   # 87 MB of rank-2 state per sample (x3 buffers): 17 GB at B = 64
    "qm9": ("qm9", 1024, 1024),
}


class Holder:
    def __init__(self, kind, hp, sd):
        self.__dict__.update(hp)
        self.model_type = kind
        self._sd = sd

    def state_dict(self):
        return self._sd

    def eval(self):
        return self


def load_workload(name):
    z = np.load(ROOT / "tests" / "golden" / f"weights_{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    keys = ["x", "adj"] + (["rank2"] if meta["is_cc"] else [])
    holders = []
    for k in keys:
        hp = meta["params"][k]
        sd = {kk[len(k) + 1:]: torch.from_numpy(z[kk]) for kk in z.files if kk.startswith(k + "/")}
        holders.append(Holder(hp["model_type"], hp, sd))
    return meta, keys, holders


def make_flags(N, B, seed=0):
    """Synthetic node-count mask: n_b uniform on [ceil(N/2), N] (SURVEY 8d fallback)."""
    rng = np.random.RandomState(seed)
    n = rng.randint((N + 1) // 2, N + 1, size=B)
    return torch.from_numpy((np.arange(N)[None, :] < n[:, None]).astype(np.float32))


def dims(meta):
    from ccsd_b200.packer import rank2_dim
    d = meta["data"]
    N, F = d["max_node_num"], d["max_feat_num"]
    if meta["is_cc"]:
        E, K = rank2_dim(N, d["d_min"], d["d_max"])
        return N, F, E, K, d["d_min"], d["d_max"]
    return N, F, 0, 0, None, None


def shapes_of(meta, B):
    N, F, E, K, _, _ = dims(meta)
    return [(B, N, F), (B, N, N)] + ([(B, E, K)] if meta["is_cc"] else [])


# ---- algorithmic FLOPs (dense-matmul 2mnk only) ------------------------------------
def mlp_flops(rows, dims_):
    return sum(2 * rows * a * b for a, b in zip(dims_[:-1], dims_[1:]))


def xa_flops(meta):
    N, F, E, K, _, _ = dims(meta)
    px, pa = meta["params"]["x"], meta["params"]["adj"]
    fl = 0
    din = F
    for _ in range(px["depth"]):
        fl += 2 * N * din * px["nhid"] + 2 * N * N * px["nhid"]
        din = px["nhid"]
    fd = F + px["depth"] * px["nhid"]
    fl += mlp_flops(N, [fd, 2 * fd, 2 * fd, F])
    L, nl, c0, ch, cf, nh, ad = (pa[k] for k in ("num_layers", "num_linears", "c_init", "c_hid", "c_final", "nhid", "adim"))
    fl += (c0 - 1) * 2 * N ** 3
    cin, kin, fdA = c0, F, c0
    for l in range(L):
        cout = ch if l < L - 1 or L == 1 else cf
        a = nh if l == 0 else ad
        fl += cin * (2 * N * kin * (2 * a + nh) + 2 * N * N * (2 * a + nh) + 2 * N * N * a)
        hid = 2 * max(cin, cout)
        fl += mlp_flops(N, [cin * nh, hid, nh])
        fl += mlp_flops(N * N, [2 * cin] + [hid] * (nl - 1) + [cout])
        cin, kin = cout, nh
        fdA += cout
    if meta["is_cc"] and pa["model_type"] == "ScoreNetworkA_CC":
        fdA += c0 + (pa["c_hid_h"] if pa["num_layers_h"] > 1 else 0) + (pa["c_final_h"] if pa["num_layers_h"] > 1 else pa["c_hid_h"])
    fl += mlp_flops(N * N, [fdA, 2 * fdA, 2 * fdA, 1])
    xa_flops.final = mlp_flops(N * N, [fdA, 2 * fdA, 2 * fdA, 1])   # the final per-edge MLP alone
    return fl


def kernel_alg_work(meta, B, pr0):
    """Algorithmic work per launch of every kernel: (flops, hbm_bytes).  FLOPs are dense-matmul 2mnk only
    (SURVEY 8d); bytes are the compulsory traffic of the rank-2 state (4 E K per sample per read or write)."""
    N, F, E, K, _, _ = dims(meta)
    x_fl, a_fl = xa_flops_split(meta)
    out = {"x_net_kernel": (B * x_fl, 0), "xa_pipeline": (B * (x_fl + a_fl), 0),
           "big_final_kernel": (B * xa_flops.final, 0), "afinal_kernel": (B * xa_flops.final, 0), "tc_afinal_kernel": (B * xa_flops.final, 0)}
    # large-graph pipeline: the aggregation GEMMs A^ (x W) of all GCN convolutions, averaged over its launches per evaluation
    px, pa = meta["params"]["x"], meta["params"]["adj"]
    L_, c0, ch, cf, nh, ad = (pa[k] for k in ("num_layers", "c_init", "c_hid", "c_final", "nhid", "adim"))
    agg = px["depth"] * 2 * N * N * px["nhid"]
    cin = c0
    for l in range(L_):
        a_ = nh if l == 0 else ad
        agg += cin * 2 * N * N * (2 * a_ + nh)
        cin = ch if l < L_ - 1 or L_ == 1 else cf
    out["big_agg_kernel"] = (B * agg / (px["depth"] + L_), 0)
    out["tc_agg_kernel"] = (B * (agg - px["depth"] * 2 * N * N * px["nhid"]) / L_, 0)
    if meta["is_cc"]:
        st = 4 * E * K * B
        out["gram_kernel"] = out["tc_gram_kernel"] = (B * 2 * E * (E + pr0) * K, st)
        # apply passes: read the state once; CORR / PRED / SCORE / EVAL also write it once (NORM does not)
        out["apply_kernel"] = out["tc_apply_kernel"] = (B * 2 * E * E * K, 2 * st)
        out["tc_apply_kernel:norm"] = (B * 2 * E * E * K, st)
        # large complexes (E > 192): K-chunked tcgen05 GEMMs (tensor bound) + an element-wise epilogue pass
        out["tc_r2big_kernel<gram>"] = (B * E * (E + pr0) * K, st)      # only the tiles that reach the diagonal: ~half of 2 E^2 K
        out["tc_r2big_kernel<hf>"] = (B * 2 * E * E * K, 2 * st)
        out["r2_epi_kernel"] = (0, 3 * st)
    return out


def xa_flops_split(meta):
    tot = xa_flops(meta)
    N, F, E, K, _, _ = dims(meta)
    px = meta["params"]["x"]
    fl, din = 0, F
    for _ in range(px["depth"]):
        fl += 2 * N * din * px["nhid"] + 2 * N * N * px["nhid"]
        din = px["nhid"]
    fd = F + px["depth"] * px["nhid"]
    fl += mlp_flops(N, [fd, 2 * fd, 2 * fd, F])
    return fl, tot - fl


# ---- whole-step algorithmic work (FLOPs per score EVALUATION of one sample, necessary computation) --
SURVEY_8D_EVAL_FLOPS = {
    "community_small": 3.916e7 / 2, "community_small_cc": 3.924e8 / 2, "qm9_cc": 1.304e7 / 2, "qm9": 2.832e6 / 2,
    "enzymes_small_cc": 4.051e7, "grid_small_cc": 2.061e11 / 2, "grid": 1.455e10 / 2,
}


# X + A_alg of the same table (the x / adj networks alone; graph-only configs: the whole evaluation)
SURVEY_8D_XA_FLOPS = {
    "community_small": 3.916e7 / 2, "community_small_cc": 3.015e6 + 2.511e7, "qm9_cc": 7.14e4 + 3.227e6, "qm9": 2.832e6 / 2,
    "enzymes_small_cc": 4.395e6 + 1.837e7, "grid_small_cc": 6.74e6 + 7.667e8, "grid": 1.455e10 / 2,
}


def step_alg_work(meta, wname, n_eval):
    """(F_alg, Bytes_alg) per sample and sampler step: SURVEY 8d's per-evaluation FLOPs x evaluations per step, and
    2 * 4 * (N F + N^2 + E K) bytes (read + write the fp32 state once per step)."""
    N, F, E, K, _, _ = dims(meta)
    if wname in SURVEY_8D_EVAL_FLOPS:
        fl, src = SURVEY_8D_EVAL_FLOPS[wname], "SURVEY 8d table"
    else:
        x_fl, a_fl = xa_flops_split(meta)
        fl, src = x_fl + a_fl + 4.0 * E * E * K, "2mnk formula (config not in the SURVEY 8d table)"
    return fl * n_eval, 2.0 * 4.0 * (N * F + N * N + E * K), src


# ---- clocks -----------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.idx)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU baseline (oracle port) ---------------------------------------------------------------
def shipped_sampler(meta, override=None):
    """(predictor, corrector) of the checkpoint's own sample_*.yaml, or the --sampler override: S4, or PC with the
    shipped predictor / corrector (Reverse + Langevin where the checkpoint ships S4)."""
    sh = meta["shipped_sampler"]
    if override == "S4":
        return "S4", "None"
    if override == "PC" and sh["predictor"] == "S4":
        return "Reverse", "Langevin"
    return sh["predictor"], sh["corrector"]


def cpu_baseline(workload, budget_s=20.0, threads=None, sampler=None, max_steps=50):
    from oracle import ccsd_oracle as O
    wname, _, B = WORKLOADS[workload]
    meta, keys, holders = load_workload(wname)
    N, F, E, K, d_min, d_max = dims(meta)
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    models = [O.Model(h.model_type, meta["params"][k], h.state_dict(), is_cc=meta["is_cc"]) for k, h in zip(keys, holders)]
    s = meta["sde"]
    sdes = [O.make_sde(s[k]["type"], s[k]["beta_min"], s[k]["beta_max"], s[k]["num_scales"]) for k in keys]
    sh = meta["shipped_sampler"]
    flags = make_flags(N, B)
    shapes = shapes_of(meta, B)

    def run(steps):
        kw = dict(snr=sh["snr"], scale_eps=sh["scale_eps"], denoise=True, eps=1e-4, d_min=d_min, d_max=d_max,
                  noise=O.NoiseSource(0), max_steps=steps)
        t0 = time.perf_counter()
        if pred == "S4":
            O.s4_solver(models, sdes, shapes, flags, **kw)
        else:
            O.pc_sampler(models, sdes, shapes, flags, predictor=pred, corrector=corr, n_steps=1, **kw)
        return time.perf_counter() - t0

    pred, corr = shipped_sampler(meta, sampler)
    t1 = run(1)  # also warms the allocator; includes prior sampling
    steps = int(max(2, min(max_steps, budget_s / max(t1, 1e-3))))
    t = run(steps)
    ms_step = 1000.0 * t / steps
    value = B / (ms_step * 1e-3 * 1000)
    return {"value": value, "unit": "complexes/s", "cores": threads, "kind": "port",
            "sample": f"{steps} sampler steps of the real 1000-step schedule at B={B} (the reference's own per-call batch), "
                      f"oracle port of ccsd/src/solver.py on torch CPU fp32, {ms_step:.1f} ms/step, extrapolated to 1000 steps",
            "ms_per_step": ms_step, "batch": B, "steps": steps, "warmup": 1}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    # steps/warmup are honoured as bounded samples: each "step" is one sampler iteration at the CPU batch
    # --steps K is honoured up to the time budget: the record carries the number of steps that were actually timed
    cb = cpu_baseline(args.workload, budget_s=float(os.environ.get("CCSD_CPU_BUDGET_S", "45")), sampler=args.sampler,
                      max_steps=max(2, args.steps))
    wname, B_gpu, _ = WORKLOADS[args.workload]
    line = {
        "impl": "reference", "metric": "sampled complexes/sec (full PC sampler)", "value": cb["value"],
        "unit": "complexes/s", "n_gpus": args.gpus, "steps": cb["steps"], "warmup": cb["warmup"],
        "steps_requested": args.steps,
        "ms_per_step": cb["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic flags + torch CPU noise; shipped-checkpoint weights",
        "config": {"workload": f"{wname} PC sampler, CPU port of the reference at B={cb['batch']}", "timed": cb["sample"]},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "complexes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0,
    }
    _emit(line)


# ---- main -------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's version banner under torchrun), so
    fd 1 is pointed at stderr for the whole run and the JSON line is written to the saved descriptor at the end."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def _emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="community_small_cc", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=None, help="per-GPU batch (default: BASELINE size)")
    ap.add_argument("--sampler", default=None, choices=["PC", "S4"], help="override the checkpoint's shipped sampler "
                    "(BASELINE configs[2] words QM9_CC with S4; its sample_qm9_CC.yaml ships Reverse + Langevin)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps", type=int, default=8)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    from ccsd_b200 import build as _build
    _build.build_cuda()  # no-op when the in-tree .so is current
    from ccsd_b200.solver import Engine, get_pc_sampler, S4_solver, quantize
    from ccsd_b200.shard import gather_rows, sharded_sample
    from ccsd_b200 import sde as bsde

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback. Use --impl reference for the CPU port.")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    wname, B, _ = WORKLOADS[args.workload]
    B = args.batch or B
    meta, keys, holders = load_workload(wname)
    N, F, E, K, d_min, d_max = dims(meta)
    s = meta["sde"]
    mk = {"VP": bsde.VPSDE, "VE": bsde.VESDE, "subVP": bsde.subVPSDE}
    sdes = [mk[s[k]["type"]](s[k]["beta_min"], s[k]["beta_max"], s[k]["num_scales"]) for k in keys]
    sh = meta["shipped_sampler"]
    pred, corr = shipped_sampler(meta, args.sampler)
    shapes = shapes_of(meta, B)
    sampler = "S4" if pred == "S4" else "PC"
    n_eval = 2 if (sampler == "PC" and corr == "Langevin") else 1
    n_total = sdes[1].N
    K_steps = max(1, min(args.steps, n_total))
    flags_host = make_flags(N, B, seed=rank).pin_memory()
    flags = flags_host.to(dev, non_blocking=True)
    ekw = dict(sampler=sampler, predictor=pred, corrector=corr, snr=sh["snr"], scale_eps=sh["scale_eps"], n_steps=1,
               denoise=True, eps=1e-4, device=dev, d_min=d_min, d_max=d_max)
    eng = Engine(holders, sdes, shapes, **ekw)
    if eng.traj_bytes() <= Engine.TRAJ_LIMIT_BYTES:
        eng.enable_traj()  # the reference records sample 0 every step (solver.py:1149-1165); so do we
    sizes = [B] * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def gather(outs):
        """single NCCL gather of the (quantised) results, BASELINE.json north_star (ccsd_b200/shard.py)"""
        if world == 1:
            return outs
        return [gather_rows(outs[0], sizes)] + [gather_rows(quantize(t), sizes) for t in outs[1:]]

    # warm-up: W untimed steps
    eng.init(flags, seed=1234, sample_offset=rank * B)
    eng.run(0, max(args.warmup, 3))
    gather(eng.read(True))
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    eng.init(flags, seed=1234, sample_offset=rank * B)
    eng.run(0, K_steps)
    outs = eng.read(True)
    gather(outs)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launches - l0
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / K_steps
    value = (B * world) / (ms_per_step * 1e-3 * n_total)

    # ---- e2e: the public API a user calls -- sampler factory (+ shard.sharded_sample on several GPUs) -- with HOST
    # buffers: pinned flags in (H2D), results quantised on the device as the reference does right after sampling
    # (sampler.py:531-543; x stays fp32) and copied to pinned host memory (D2H), everything inside the timed region.
    # On N GPUs every rank copies ITS shard of the gathered result (the downstream graph conversion is per-sample CPU
    # work of that rank's process).
    fac = S4_solver if sampler == "S4" else get_pc_sampler
    kw = dict(predictor=pred, corrector=corr, snr=sh["snr"], scale_eps=sh["scale_eps"], n_steps=1,
              probability_flow=False, continuous=True, denoise=True, eps=1e-4, device=dev)
    del eng
    torch.cuda.empty_cache()

    def make_sampler(b):
        k2 = dict(kw)
        sh3 = shapes_of(meta, b)
        if meta["is_cc"]:
            k2.update(is_cc=True, sde_rank2=sdes[2], shape_rank2=sh3[2], d_min=d_min, d_max=d_max)
        return fac(sdes[0], sdes[1], sh3[0], sh3[1], **k2)

    cache = {}

    def sampler_of(b):   # one plan per shard size (the factory's own cache lives in the returned callable)
        if b not in cache:
            cache[b] = make_sampler(b)
        return cache[b]

    host_out = [torch.empty(shapes[0], dtype=torch.float32).pin_memory()] + \
               [torch.empty(sh_, dtype=torch.uint8).pin_memory() for sh_ in shapes[1:]]
    all_flags_host = torch.cat([make_flags(N, B, seed=r) for r in range(world)]).pin_memory() if world > 1 else flags_host

    def e2e_call(seed, steps):
        fl = all_flags_host.to(dev, non_blocking=True)                                    # H2D
        if world > 1:
            res = sharded_sample(sampler_of, holders, fl, seed=seed, quantize_fn=quantize, max_steps=steps, record_traj=True)
            res = [t[rank * B:(rank + 1) * B] for t in res]
        else:
            out = sampler_of(B)(*holders, fl, seed=seed, sample_offset=0, max_steps=steps)
            res = [out[0]] + [quantize(t) for t in out[1:len(shapes)]]
        for h, t in zip(host_out, res):
            h.copy_(t, non_blocking=True)                                                 # D2H
        return res

    e2e_call(1, 3)   # builds the plan
    barrier()
    t0 = time.perf_counter()
    e2e_call(1234, K_steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = (B * world) / (e2e_s / K_steps * n_total)
    h2d = all_flags_host.numel() * 4
    d2h = sum(h.numel() * h.element_size() for h in host_out)
    cache.clear()
    torch.cuda.empty_cache()

    # ---- per-kernel device times (CUDA events on the launching stream) + roofline of the dominant unit and of the step
    roofline = None
    prof_summary = {}
    if rank == 0:
        eng2 = Engine(holders, sdes, shapes, **ekw)
        eng2.init(flags, seed=1234, sample_offset=0)
        eng2.run(0, 3)
        torch.cuda.synchronize()
        eng2.set_profiling(True)
        eng2.run(3, 3 + args.profile_steps)
        recs = eng2.get_profile()
        eng2.set_profiling(False)
        for nm, t in recs:
            a = prof_summary.setdefault(nm, [0.0, 0])
            a[0] += t
            a[1] += 1
        tot = sum(v[0] for v in prof_summary.values()) or 1.0
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        src = "measured (MEASURED_PEAKS.json, sustained)"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        if not pk.exists():
            src = "fallback (B200_PROFILING.md)"
        work = kernel_alg_work(meta, B, eng2.desc.neta.n_proj_rows[0] if meta["is_cc"] else 0)
        peak_hbm = peaks.get("hbm_gbs", 6650.0)
        XA = ("x_net_kernel", "tc_xfin_kernel", "attn_channel_kernel", "tc_attn_kernel", "attn_finish_kernel", "tc_edge_kernel", "proj1_kernel", "hodge_kernel",
              "hodge_base_kernel", "afinal_kernel", "tc_afinal_kernel", "big_prep_kernel", "big_pow_kernel", "big_deg_kernel", "big_xw_kernel",
              "big_agg_kernel", "tc_agg_kernel", "big_attn_kernel", "big_node_kernel", "big_edge_kernel", "big_edge_pair_kernel",
              "big_mirror_kernel", "big_final_kernel", "big_xfin_kernel")
        R2 = ("tc_gram_kernel", "gram_kernel", "tc_apply_kernel", "apply_kernel", "tc_r2big_kernel<gram>", "tc_r2big_kernel<hf>", "r2_epi_kernel", "tc_hnorm_kernel", "znorm_kernel")
        kern = {}
        for k_, v in prof_summary.items():
            ms_l = v[0] / v[1]
            fl, by = work.get(k_, (0, 0))
            kern[k_] = {"ms_per_launch": ms_l, "launches": v[1], "share": v[0] / tot,
                        "alg_tflops": fl / (ms_l * 1e-3) / 1e12, "alg_gbs": by / (ms_l * 1e-3) / 1e9}
        # the x / adj network pipeline as ONE unit of work (its kernels implement one score evaluation); likewise the
        # rank-2 passes (Gram + apply passes of one step)
        xa_ms = sum(prof_summary[k_][0] for k_ in XA if k_ in prof_summary)
        r2_ms = sum(prof_summary[k_][0] for k_ in R2 if k_ in prof_summary)
        n_evals = n_eval * args.profile_steps
        x_fl, a_fl = xa_flops_split(meta)
        xa_alg = SURVEY_8D_XA_FLOPS.get(wname, x_fl + a_fl)
        if xa_ms:
            kern["xa_pipeline"] = {"ms_per_launch": xa_ms / n_evals, "launches": n_evals, "share": xa_ms / tot,
                                   "alg_tflops": B * xa_alg / (xa_ms / n_evals * 1e-3) / 1e12, "alg_gbs": 0.0,
                                   "kernels": [k_ for k_ in XA if k_ in prof_summary]}
        Fs, Bs, wsrc = step_alg_work(meta, wname, n_eval)
        if r2_ms:
            kern["rank2_passes"] = {"ms_per_launch": r2_ms / args.profile_steps, "launches": args.profile_steps, "share": r2_ms / tot,
                                    "alg_tflops": B * n_eval * 4.0 * E * E * K / (r2_ms / args.profile_steps * 1e-3) / 1e12,
                                    "alg_gbs": B * 8.0 * E * K / (r2_ms / args.profile_steps * 1e-3) / 1e9,
                                    "kernels": [k_ for k_ in R2 if k_ in prof_summary]}
        # dominant UNIT by aggregated time: the x/adj pipeline (tensor roofline) or the rank-2 passes (HBM roofline)
        traffic = None
        tf = ROOT / "profiles" / "ncu_traffic.json"
        if tf.exists():
            traffic = json.loads(tf.read_text()).get(wname, {})
        if r2_ms >= xa_ms:
            dom = max((k_ for k_ in R2 if k_ in prof_summary), key=lambda k_: prof_summary[k_][0])
            avg_ms = prof_summary[dom][0] / prof_summary[dom][1]
            fl, by = work.get(dom, (0, 0))
            if dom == "tc_apply_kernel" and sampler == "PC" and corr == "Langevin" and prof_summary[dom][1] >= 3 * args.profile_steps:
                # NORM (read only), CORR, PRED passes.  With the Langevin norms from Gram quantities (tc_hnorm) there is no NORM
                # pass: both launches of a step read and write the state
                by = (work["tc_apply_kernel:norm"][1] + 2 * work["tc_apply_kernel"][1]) / 3.0
            if dom.startswith("tc_r2big_kernel"):   # real GEMMs: tensor roofline (algorithmic 2mnk; bf16x3 executes three MMAs per product)
                ach = fl / (avg_ms * 1e-3) / 1e12
                roofline = {"unit_of_work": "rank2_passes", "kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s",
                            "frac": ach / peak_tf, "traffic": (traffic or {}).get(dom), "peak_source": src, "avg_launch_ms": avg_ms,
                            "algorithmic_flops_per_launch": fl, "share_of_step": prof_summary[dom][0] / tot, "unit_share_of_step": r2_ms / tot}
            else:
                ach = by / (avg_ms * 1e-3) / 1e9
                roofline = {"unit_of_work": "rank2_passes", "kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak_hbm, "unit": "GB/s",
                            "frac": ach / peak_hbm, "traffic": (traffic or {}).get(dom), "peak_source": src.replace("sustained", "hbm_gbs"),
                            "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": by, "share_of_step": prof_summary[dom][0] / tot,
                            "unit_share_of_step": r2_ms / tot}
        else:
            avg_ms = xa_ms / n_evals
            ach = B * xa_alg / (avg_ms * 1e-3) / 1e12
            roofline = {"unit_of_work": "xa_pipeline", "kernel": "xa_pipeline (all kernels of one score evaluation of the x / adj networks)",
                        "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf, "traffic": None,
                        "peak_source": src, "avg_launch_ms": avg_ms, "algorithmic_flops_per_launch": B * xa_alg,
                        "share_of_step": xa_ms / tot, "unit_share_of_step": xa_ms / tot}
        # the WHOLE step against both rooflines (F_alg and Bytes_alg per sample and step)
        step_s = ms_per_step * 1e-3
        roofline["step_frac_tensor"] = B * Fs / step_s / (peak_tf * 1e12)
        roofline["step_frac_hbm"] = B * Bs / step_s / (peak_hbm * 1e9)
        roofline["step_alg"] = {"flops_per_sample_step": Fs, "bytes_per_sample_step": Bs, "source": wsrc}
        roofline["kernels"] = kern
        del eng2

    if rank == 0:
        line = {
            "metric": "sampled complexes/sec (full PC sampler)", "value": value, "unit": "complexes/s",
            "n_gpus": world, "steps": K_steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "us_per_score_step": ms_per_step * 1000.0, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic node-count flags + Philox noise; shipped-checkpoint weights (tests/golden)",
            "config": {"workload": f"{wname}: N={N} F={F} E={E} K={K}, {sampler} sampler {pred}+{corr} "
                                   f"snr={sh['snr']} scale_eps={sh['scale_eps']}, {n_total}-step schedule, batch {B} per GPU",
                       "global_batch": B * world, "parallelism": f"dp{world} (batch shards, no per-step collective, one NCCL gather)",
                       "l2": ("state per step (rank-2 tensors) is larger than L2; no flush needed" if meta["is_cc"] else
                              "graph-only: the state is L2 resident by design (compute-bound pipeline); intermediates stream through L2/HBM"),
                       "timed": f"{K_steps} sampler steps incl. prior sampling" + (" = the whole sampler run" if K_steps == n_total else " (scaled to 1000)")},
            "clocks": clk, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value, "unit": "complexes/s", "h2d_bytes_per_step": h2d / K_steps, "d2h_bytes_per_step": d2h / K_steps,
                    "seconds": e2e_s,
                    "call": ("ccsd_b200.get_pc_sampler / S4_solver(...)(models, init_flags)" + (" through ccsd_b200.shard.sharded_sample" if world > 1 else "") +
                             ": pinned host flags in, ccsd_b200.quantize of adj / rank2 on the device (sampler.py:531-543), x fp32 + "
                             "adj / rank2 uint8 to pinned host memory")},
            "roofline": roofline,
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(args.workload, sampler=args.sampler).items() if k in ("value", "unit", "cores", "kind", "sample")}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
