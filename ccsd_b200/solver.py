"""Drop-in replacements for ``ccsd.src.solver.get_pc_sampler`` / ``S4_solver``
(ccsd/src/solver.py:856-1176, 1179-1563): same factory signatures, same returned-callable contract,
same exceptions -- the work runs in the sm_100a kernels behind the C ABI (include/ccsd_b200.h).

Returned callable:
    graph: ``fn(model_x, model_adj, init_flags) -> (x, adj, n_evals, diff_traj)``
    CC:    ``fn(model_x, model_adj, model_rank2, init_flags) -> (x, adj, rank2, n_evals, diff_traj)``
Keyword-only extras (not in the reference, all optional): ``seed`` (Philox seed; default: drawn
from torch's global generator so ``load_seed`` keeps runs reproducible), ``sample_offset`` (global
index of this shard's first sample), ``noise`` (an injected raw-normal stream for parity tests),
``max_steps`` (stop early on the real schedule), ``record_traj``.

There is no CPU path: a CUDA device and the built extension are required.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat
from . import packer
from .schedule import build_schedule
from .sde import sde_kind


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class InjectedNoise:
    """Raw standard-normal draws in the reference's order, for parity runs.

    prior: list [x (B,N,F), adj (B,N,N)[, rank2 (B,E,K)]] of RAW draws (before symmetrise / mask).
    steps: per step, list per object of tensors [n_draws, B, ...] (PC+Langevin: corrector,
    predictor; PC+None: predictor; S4: correction, first half, second half).
    """

    def __init__(self, prior: Sequence[torch.Tensor], steps: Sequence[Sequence[torch.Tensor]]):
        self.prior, self.steps = list(prior), [list(s) for s in steps]

    @staticmethod
    def from_flat_log(log: Sequence[torch.Tensor], n_obj: int, n_draws: int, n_steps: int,
                      n_lang: Optional[int] = None) -> "InjectedNoise":
        """Regroup a flat draw log in reference order: prior (one draw per object), then per step the draws in
        the order solver.py consumes torch's generator.  S4 / PC with ``n_lang`` in (None, 1): ``n_draws`` rounds
        of one draw per object.  PC + Langevin with ``n_lang`` inner steps: every corrector runs its whole loop
        before the next object's (solver.py:692, 760, 1123-1140) -- x's n_lang draws, adj's, rank2's -- then one
        predictor round."""
        prior = list(log[:n_obj])
        steps, pos = [], n_obj
        for _ in range(n_steps):
            chunk = log[pos:pos + n_draws * n_obj]
            pos += n_draws * n_obj
            if n_lang is None or n_lang <= 1:
                steps.append([torch.stack([chunk[dr * n_obj + k] for dr in range(n_draws)]) for k in range(n_obj)])
            else:
                assert n_draws == n_lang + 1
                steps.append([torch.stack([chunk[k * n_lang + s] for s in range(n_lang)] + [chunk[n_obj * n_lang + k]])
                              for k in range(n_obj)])
        return InjectedNoise(prior, steps)


class Engine:
    """One bound plan: weights blob, schedule, workspace; thin wrapper over the C ABI."""

    def __init__(
        self, models: Sequence[Any], sdes: Sequence[Any], shapes: Sequence[Sequence[int]], *, sampler: str,
        predictor: str = "Euler", corrector: str = "None", snr: float = 0.1, scale_eps: float = 1.0, n_steps: int = 1,
        probability_flow: bool = False, denoise: bool = True, eps: float = 1e-3, device="cuda",
        d_min: Optional[int] = None, d_max: Optional[int] = None, nets: Optional[int] = None,
    ) -> None:
        self.lib = nat.load()
        self.device = torch.device(device)
        if self.device.type != "cuda" and not nat.is_emulation():
            raise RuntimeError("ccsd_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        if self.device.type == "cuda" and nat.is_emulation():
            raise RuntimeError("the host-emulation build cannot drive a CUDA device")
        self.is_cc = len(shapes) == 3
        B, N, F = (int(v) for v in shapes[0])
        if tuple(shapes[1]) != (B, N, N):
            raise ValueError(f"shape_adj {tuple(shapes[1])} does not match shape_x {tuple(shapes[0])}")
        d = nat.PlanDesc()
        d.B, d.N, d.F, d.is_cc = B, N, F, int(self.is_cc)
        if self.is_cc:
            E, K = packer.rank2_dim(N, d_min, d_max)
            if tuple(shapes[2]) != (B, E, K):
                raise ValueError(f"shape_rank2 {tuple(shapes[2])} != (B, {E}, {K}) for N={N}, d_min={d_min}, d_max={d_max}")
            d.E, d.K, d.d_min, d.d_max = E, K, int(d_min), int(d_max)
        if sampler not in ("PC", "S4"):
            raise ValueError(sampler)
        d.sampler = nat.SAMPLER_PC if sampler == "PC" else nat.SAMPLER_S4
        if sampler == "PC":
            if predictor not in ("Reverse", "Euler"):
                raise NotImplementedError(f"Predictor {predictor} not yet supported. Select from [Reverse, Euler].")
            if corrector not in ("Langevin", "None"):
                raise NotImplementedError(f"Corrector {corrector} not yet supported. Select from [Langevin, None].")
        d.use_corrector = int(sampler == "PC" and corrector == "Langevin")
        d.n_lang_steps = int(n_steps) if d.use_corrector else 1
        d.denoise = int(bool(denoise))
        d.n_diff_steps = int(sdes[1].N)
        d.snr, d.scale_eps = float(snr), float(scale_eps)
        present, blob_host = self._pack(d, models)
        d.nets = present if nets is None else nets
        self.desc = d
        self.n_obj = 3 if self.is_cc else 2
        self.n_draws = 3 if sampler == "S4" else (d.n_lang_steps + 1 if d.use_corrector else 1)
        self.sizes = [B * N * F, B * N * N] + ([B * d.E * d.K] if self.is_cc else [])
        self.shapes = [tuple(int(v) for v in s) for s in shapes]
        sched = build_schedule(list(sdes), sampler=sampler, predictor=predictor, probability_flow=probability_flow, eps=eps)
        self.schedule = np.ascontiguousarray(sched, np.float32)
        self._blob_host = blob_host
        self._nets_arg = nets
        with self._guard():
            self.weights = torch.from_numpy(blob_host).to(self.device)
            handle = C.c_void_p()
            nat.check(self.lib.ccsd_plan_create(
                C.byref(d), self.schedule.ctypes.data_as(C.c_void_p), self.weights.data_ptr(), self.weights.numel(), C.byref(handle)))
            self.handle = handle
            nbytes = int(self.lib.ccsd_plan_workspace_bytes(handle))
            self.workspace = torch.empty(nbytes + 256, dtype=torch.uint8, device=self.device)
            base = self.workspace.data_ptr()
            self._ws_ptr = (base + 255) // 256 * 256
            nat.check(self.lib.ccsd_plan_bind(handle, self._ws_ptr, nbytes, self._stream()))
        self.traj: Optional[List[torch.Tensor]] = None

    # -- helpers --
    def _pack(self, d, models):
        """Fill the topology part of descriptor `d` and return (present-network bits, packed fp32 blob)."""
        blob = packer.Blob()
        present = 0
        mx, ma = models[0], models[1]
        mf = models[2] if self.is_cc and len(models) > 2 else None
        if mx is not None:
            packer.pack_netx(d.netx, blob, mx)
            present |= 1
        if ma is not None:
            packer.pack_neta(d.neta, blob, ma, d.K)
            present |= 2
        if mf is not None:
            packer.pack_netf(d.netf, blob, mf)
            present |= 4
        return present, blob.finish()

    def refresh_weights(self, models: Sequence[Any]) -> bool:
        """Re-read the models' CURRENT parameters (the reference reads the live modules on every call: an EMA
        ``copy_to``, ``load_state_dict`` or optimizer step between two calls must be seen).  Returns False when the
        topology changed, i.e. the plan has to be rebuilt; otherwise uploads the blob if any weight differs."""
        d2 = nat.PlanDesc()
        C.memmove(C.byref(d2), C.byref(self.desc), C.sizeof(nat.PlanDesc))
        C.memset(C.byref(d2.netx), 0, C.sizeof(d2.netx)); C.memset(C.byref(d2.neta), 0, C.sizeof(d2.neta))
        C.memset(C.byref(d2.netf), 0, C.sizeof(d2.netf))
        present, blob = self._pack(d2, models)
        d2.nets = present if self._nets_arg is None else self._nets_arg
        if bytes(d2) != bytes(self.desc) or blob.shape != self._blob_host.shape:
            return False
        if not np.array_equal(blob, self._blob_host):
            # folded constants of the affine ScoreNetworkF live in the descriptor: covered by the bytes() comparison above
            self._blob_host = blob
            with self._guard():
                self.weights.copy_(torch.from_numpy(blob), non_blocking=False)
        return True

    def _guard(self):
        """Every ABI call launches on the CURRENT device: make it the engine's (a plan on cuda:1 must not launch on cuda:0)."""
        if self.device.type == "cuda":
            return torch.cuda.device(self.device)
        import contextlib
        return contextlib.nullcontext()

    def _stream(self):
        if self.device.type == "cuda":
            return torch.cuda.current_stream(self.device).cuda_stream
        return None

    def _dev(self, t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if t is None:
            return None
        return t.to(device=self.device, dtype=torch.float32).contiguous()

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ccsd_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- API --
    TRAJ_LIMIT_BYTES = 16 << 30   # diff_traj of sample 0 on the device: grid_small_CC would need 87 GB

    def traj_bytes(self) -> int:
        d = self.desc
        return 4 * d.n_diff_steps * (d.N * d.F + d.N * d.N + (d.E * d.K if self.is_cc else 0))

    def enable_traj(self) -> None:
        n = self.desc.n_diff_steps
        d = self.desc
        self.traj = [torch.zeros(n, d.N, d.F, device=self.device), torch.zeros(n, d.N, d.N, device=self.device)]
        if self.is_cc:
            self.traj.append(torch.zeros(n, d.E, d.K, device=self.device))
        nat.check(self.lib.ccsd_plan_set_traj(
            self.handle, _ptr(self.traj[0]), _ptr(self.traj[1]), _ptr(self.traj[2]) if self.is_cc else None))

    def disable_traj(self) -> None:
        self.traj = None
        nat.check(self.lib.ccsd_plan_set_traj(self.handle, None, None, None))

    def init(self, flags: torch.Tensor, prior: Optional[Sequence[torch.Tensor]] = None, seed: int = 0,
             sample_offset: int = 0) -> None:
        d = self.desc
        if tuple(flags.shape) != (d.B, d.N):
            raise ValueError(f"init_flags shape {tuple(flags.shape)} != ({d.B}, {d.N})")
        fl = self._dev(flags)
        pr = [self._dev(p) for p in prior] if prior is not None else [None] * 3
        while len(pr) < 3:
            pr.append(None)
        for p, s in zip(pr, self.shapes):
            if p is not None and tuple(p.shape) != s:
                raise ValueError(f"prior shape {tuple(p.shape)} != {s}")
        self._keep = (fl, pr)
        with self._guard():
            nat.check(self.lib.ccsd_plan_init(self.handle, _ptr(fl), _ptr(pr[0]), _ptr(pr[1]), _ptr(pr[2]),
                                              C.c_uint64(seed & (2 ** 64 - 1)), C.c_int64(sample_offset), self._stream()))

    def step(self, i: int, noise: Optional[Sequence[torch.Tensor]] = None) -> None:
        nz = [None] * 3
        if noise is not None:
            for k in range(self.n_obj):
                t = self._dev(noise[k])
                if t.numel() != self.n_draws * self.sizes[k]:
                    raise ValueError(f"noise[{k}] has {t.numel()} elements, expected {self.n_draws} x {self.sizes[k]}")
                nz[k] = t
        self._keep_noise = nz
        with self._guard():
            nat.check(self.lib.ccsd_plan_step(self.handle, int(i), _ptr(nz[0]), _ptr(nz[1]), _ptr(nz[2]), self._stream()))

    def run(self, begin: int, end: int) -> None:
        with self._guard():
            nat.check(self.lib.ccsd_plan_run(self.handle, int(begin), int(end), self._stream()))

    def read(self, want_mean: bool) -> List[torch.Tensor]:
        outs = [torch.empty(s, dtype=torch.float32, device=self.device) for s in self.shapes]
        with self._guard():
            nat.check(self.lib.ccsd_plan_read(self.handle, int(want_mean), _ptr(outs[0]), _ptr(outs[1]),
                                              _ptr(outs[2]) if self.is_cc else None, self._stream()))
        return outs

    def score(self, which: int, x: torch.Tensor, adj: torch.Tensor, rank2: Optional[torch.Tensor],
              flags: Optional[torch.Tensor]) -> torch.Tensor:
        """Raw network output model(x, adj[, rank2], flags) -- the per-step parity seam."""
        d = self.desc
        if flags is None:
            flags = torch.ones(d.B, d.N)
        x, adj, rank2, flags = self._dev(x), self._dev(adj), self._dev(rank2), self._dev(flags)
        if tuple(x.shape) != self.shapes[0] or tuple(adj.shape) != self.shapes[1]:
            raise ValueError("score(): input shapes do not match the plan")
        out = torch.empty(self.shapes[which], dtype=torch.float32, device=self.device)
        with self._guard():
            nat.check(self.lib.ccsd_score_eval(self.handle, int(which), _ptr(x), _ptr(adj), _ptr(rank2), _ptr(flags),
                                               _ptr(out), self._stream()))
        return out

    def debug_gram(self, rank2: torch.Tensor, use_tc: bool):
        """(H, P0) of `rank2` through the fp32 FMA or the tcgen05 Gram kernel (test seam)."""
        d = self.desc
        r2 = self._dev(rank2)
        pr0 = int(self.lib.ccsd_plan_info(self.handle, 12))   # Gram projection columns (hodge layer 0 [+ folded layer 1])
        H = torch.empty(d.B, d.E, d.E, dtype=torch.float32, device=self.device)
        P0 = torch.empty(d.B, d.E, max(pr0, 1), dtype=torch.float32, device=self.device)
        with self._guard():
            nat.check(self.lib.ccsd_debug_gram(self.handle, _ptr(r2), _ptr(H), _ptr(P0) if pr0 else None, int(use_tc),
                                               self._stream()))
        return H, P0[:, :, :pr0]

    def set_profiling(self, on: bool) -> None:
        nat.check(self.lib.ccsd_plan_set_profiling(self.handle, int(on)))

    def get_profile(self, max_records: int = 65536):
        """[(kernel name, ms)] for every launch since set_profiling(True) (waits for the events)."""
        names = C.create_string_buffer(max_records * 32)
        ms = (C.c_float * max_records)()
        n = self.lib.ccsd_plan_get_profile(self.handle, max_records, names, 32, ms)
        if n < 0:
            nat.check(n)
        raw = names.raw
        return [(raw[i * 32:(i + 1) * 32].split(b"\0")[0].decode(), float(ms[i])) for i in range(n)]

    def info(self) -> dict:
        names = ("xa_smem_bytes", "xa_threads", "apply_smem_bytes", "tc_gram", "tc_apply", "f_mode", "fin_rows",
                 "smem_x_net", "smem_attn_channel", "smem_attn_finish", "smem_hodge", "smem_afinal")
        return {n: int(self.lib.ccsd_plan_info(self.handle, i)) for i, n in enumerate(names)}

    @property
    def launches(self) -> int:
        return int(self.lib.ccsd_plan_launch_count(self.handle))


def quantize(t: torch.Tensor, thr: float = 0.5, mol: bool = False) -> torch.Tensor:
    """Device quantize / quantize_mol (ccsd/src/utils/graph_utils.py:181-213) -> uint8."""
    lib = nat.load()
    t = t.contiguous().to(torch.float32)
    out = torch.empty(t.shape, dtype=torch.uint8, device=t.device)
    stream = torch.cuda.current_stream(t.device).cuda_stream if t.device.type == "cuda" else None
    if t.device.type != "cuda" and not nat.is_emulation():
        raise RuntimeError("ccsd_b200 needs a CUDA device; there is no CPU fallback")
    nat.check(lib.ccsd_quantize(t.data_ptr(), out.data_ptr(), t.numel(), float(thr), int(mol), stream))
    return out


def mol_onehot(x: torch.Tensor, adj: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """The tensor post-processing between the sampler and ``gen_mol`` in ``Sampler_mol.sample``
    (ccsd/src/sampler.py:814-825) as one device pass: returns ``(x [B,N,F+1] int64, adj [B,4,N,N] int64)``."""
    lib = nat.load()
    x = x.contiguous().to(torch.float32)
    adj = adj.contiguous().to(torch.float32)
    if x.dim() != 3 or adj.dim() != 3 or adj.shape[0] != x.shape[0] or adj.shape[1] != x.shape[1] or adj.shape[2] != x.shape[1]:
        raise ValueError("mol_onehot: x must be [B,N,F] and adj [B,N,N]")
    if x.device.type != "cuda" and not nat.is_emulation():
        raise RuntimeError("ccsd_b200 needs a CUDA device; there is no CPU fallback")
    B, N, F = x.shape
    x_out = torch.empty(B, N, F + 1, dtype=torch.int64, device=x.device)
    adj_out = torch.empty(B, 4, N, N, dtype=torch.int64, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream if x.device.type == "cuda" else None
    nat.check(lib.ccsd_mol_onehot(x.data_ptr(), adj.data_ptr(), x_out.data_ptr(), adj_out.data_ptr(), B, N, F, stream))
    return x_out, adj_out


# ---------------------------------------------------------------------------------------------
def _check_sdes(sdes, continuous: bool) -> None:
    for s in sdes:
        sde_kind(s)  # NotImplementedError for unknown classes (losses.py:101-102)
    if not continuous:
        raise NotImplementedError("Discrete not supported")  # losses.py:69, 98


def _set_eval(models) -> None:
    for m in models:
        if hasattr(m, "eval"):
            m.eval()  # get_score_fn(train=False) side effect (losses.py:38-39)


def _make_sampler(
    sampler: str, sde_x, sde_adj, shape_x, shape_adj, predictor, corrector, snr, scale_eps, n_steps, probability_flow,
    continuous, denoise, eps, device, is_cc, sde_rank2, shape_rank2, d_min, d_max,
) -> Callable:
    if sampler == "PC":
        if predictor not in ("Reverse", "Euler"):
            raise NotImplementedError(f"Predictor {predictor} not yet supported. Select from [Reverse, Euler].")
        if corrector not in ("Langevin", "None"):
            raise NotImplementedError(f"Corrector {corrector} not yet supported. Select from [Langevin, None].")
    sdes = [sde_x, sde_adj] + ([sde_rank2] if is_cc else [])
    shapes = [tuple(shape_x), tuple(shape_adj)] + ([tuple(shape_rank2)] if is_cc else [])
    cache: Dict[str, Any] = {}

    def run(models: Sequence[Any], init_flags: torch.Tensor, *, seed: Optional[int] = None, sample_offset: int = 0,
            noise: Optional[InjectedNoise] = None, max_steps: Optional[int] = None, record_traj: bool = True):
        _check_sdes(sdes, continuous)
        _set_eval(models)
        # One cached plan per factory call.  The key holds the model objects themselves (so an id cannot be reused by
        # another object) and the weights are re-read from the live modules on EVERY call, as the reference does.
        eng = None
        ent = cache.get("engine")
        if ent is not None and len(ent[0]) == len(models) and all(a is b for a, b in zip(ent[0], models)):
            if ent[1].refresh_weights(models):
                eng = ent[1]
        if eng is None:
            eng = Engine(models, sdes, shapes, sampler=sampler, predictor=predictor, corrector=corrector, snr=snr,
                         scale_eps=scale_eps, n_steps=n_steps, probability_flow=probability_flow, denoise=denoise,
                         eps=eps, device=device, d_min=d_min, d_max=d_max)
            cache["engine"] = (list(models), eng)
        n = eng.desc.n_diff_steps
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        if record_traj and eng.traj_bytes() > Engine.TRAJ_LIMIT_BYTES:
            # the reference appends sample 0 of every object at every step (solver.py:1149-1165); at E x K = 2e7 that
            # list does not fit any memory -- return an empty trajectory instead of failing
            import warnings
            warnings.warn(f"diff_traj would need {eng.traj_bytes() / 2 ** 30:.0f} GiB; not recorded")
            record_traj = False
        if record_traj and eng.traj is None:
            eng.enable_traj()
        elif not record_traj and eng.traj is not None:
            eng.disable_traj()
        eng.init(init_flags, prior=noise.prior if noise is not None else None, seed=seed, sample_offset=sample_offset)
        steps = n if max_steps is None else min(n, max_steps)
        if noise is not None:
            for i in range(steps):
                eng.step(i, noise.steps[i])
        else:
            eng.run(0, steps)
        outs = eng.read(want_mean=bool(denoise))
        if record_traj:
            # detached copies (the reference appends clones, solver.py:987-995): a later call of this sampler reuses the
            # engine's buffers; the reference's list has one entry per executed step
            snap = [t[:steps].clone() for t in eng.traj]
            diff_traj = [[t[i] for t in snap] for i in range(steps)]
        else:
            diff_traj = []
        n_evals = n * (n_steps + 1) if sampler == "PC" else 0  # solver.py:1001, 1172, 1369, 1559
        return (*outs, n_evals, diff_traj)

    if not is_cc:
        def sampler_fn(model_x, model_adj, init_flags, **kw):
            return run([model_x, model_adj], init_flags, **kw)
    else:
        def sampler_fn(model_x, model_adj, model_rank2, init_flags, **kw):
            return run([model_x, model_adj, model_rank2], init_flags, **kw)
    sampler_fn.__name__ = "pc_sampler" if sampler == "PC" else "s4_solver"
    return sampler_fn


def get_pc_sampler(
    sde_x, sde_adj, shape_x: Sequence[int], shape_adj: Sequence[int], predictor: str = "Euler",
    corrector: str = "None", snr: float = 0.1, scale_eps: float = 1.0, n_steps: int = 1,
    probability_flow: bool = False, continuous: bool = False, denoise: bool = True, eps: float = 1e-3,
    device: str = "cuda", is_cc: bool = False, sde_rank2=None, shape_rank2: Optional[Sequence[int]] = None,
    d_min: Optional[int] = None, d_max: Optional[int] = None,
) -> Callable:
    """ccsd/src/solver.py:856-875 (signature) / :917-1176 (behaviour)."""
    return _make_sampler("PC", sde_x, sde_adj, shape_x, shape_adj, predictor, corrector, snr, scale_eps, n_steps,
                         probability_flow, continuous, denoise, eps, device, is_cc, sde_rank2, shape_rank2, d_min, d_max)


def S4_solver(
    sde_x, sde_adj, shape_x: Sequence[int], shape_adj: Sequence[int], predictor: str = "None",
    corrector: str = "None", snr: float = 0.1, scale_eps: float = 1.0, n_steps: int = 1,
    probability_flow: bool = False, continuous: bool = False, denoise: bool = True, eps: float = 1e-3,
    device: str = "cuda", is_cc: bool = False, sde_rank2=None, shape_rank2: Optional[Sequence[int]] = None,
    d_min: Optional[int] = None, d_max: Optional[int] = None,
) -> Callable:
    """ccsd/src/solver.py:1179-1198 (signature) / :1240-1563 (behaviour).  predictor, corrector,
    n_steps and probability_flow are accepted and ignored, as in the reference."""
    return _make_sampler("S4", sde_x, sde_adj, shape_x, shape_adj, predictor, corrector, snr, scale_eps, n_steps,
                         probability_flow, continuous, denoise, eps, device, is_cc, sde_rank2, shape_rank2, d_min, d_max)
