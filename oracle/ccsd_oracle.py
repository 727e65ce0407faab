"""CPU oracle for the CCSD reverse-SDE sampler hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a vectorised fp32 restatement (torch CPU tensors, no autograd) of the
reference algorithm: the score networks, the SDE coefficient maths and the PC / S4
sampling loops.  It is the checker the CUDA path is compared against.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / ``--impl reference`` legs may
import it; nothing under ``ccsd_b200/`` does, and the product path raises when the CUDA
extension is missing.

Why torch-CPU and not C/numpy: the reference itself is torch fp32 and the parity bar is
"the reference fp32 path" (BASELINE.json north_star).  Using the same aten CPU matmuls keeps
the restatement within ~1e-6 of the imported reference (checked by
``oracle/validate_against_reference.py`` in the build container and pinned by the golden
fixtures in ``tests/golden/``; the reference's own known-answer tests are replayed in
``tests/test_oracle_golden.py``).

Every function cites the reference file:line it restates (paths relative to
/root/reference/).  No reference source is copied: the reference is class/Module based,
this is a functional restatement over plain ``state_dict`` dictionaries with closed-form
masks instead of the reference's per-node Python loops.
"""
from __future__ import annotations

import math
from itertools import combinations
from functools import lru_cache
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as Fnn

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------------------
# combinatorics / masks
# --------------------------------------------------------------------------------------
def rank2_dim(N: int, d_min: int, d_max: int) -> Tuple[int, int]:
    """(E, K) = (N choose 2, sum_d N choose d).  ccsd/src/utils/cc_utils.py:268-283."""
    return (N * (N - 1)) // 2, sum(math.comb(N, d) for d in range(d_min, d_max + 1))


@lru_cache(maxsize=16)
def cell_tables(N: int, d_min: int, d_max: int) -> Tuple[Tensor, Tensor]:
    """Membership matrices: edges (E x N) and rank-2 cells (K x N), 0/1 float.

    Edge order = combinations(range(N), 2); cell order = combinations(range(N), d) for
    d = d_min..d_max concatenated (cc_utils.py:72-96, get_cells).
    """
    E, K = rank2_dim(N, d_min, d_max)
    em = torch.zeros(E, N)
    for e, (i, j) in enumerate(combinations(range(N), 2)):
        em[e, i] = 1.0
        em[e, j] = 1.0
    cm = torch.zeros(K, N)
    k = 0
    for d in range(d_min, d_max + 1):
        for c in combinations(range(N), d):
            cm[k, list(c)] = 1.0
            k += 1
    return em, cm


def mask_x(x: Tensor, flags: Optional[Tensor]) -> Tensor:
    """graph_utils.py:25-37."""
    if flags is None:
        return x
    return x * flags[:, :, None]


def mask_adjs(adjs: Tensor, flags: Optional[Tensor]) -> Tensor:
    """graph_utils.py:40-59 (3-D and 4-D)."""
    if flags is None:
        return adjs
    f = flags.unsqueeze(1) if adjs.dim() == 4 else flags
    return adjs * f.unsqueeze(-1) * f.unsqueeze(-2)


def edge_flags(flags: Tensor) -> Tensor:
    """flags_left of get_rank2_flags / get_hodge_adj_flags in closed form:
    an edge (i,j) survives iff neither node has flag 0.  cc_utils.py:527-557, 1591-1612."""
    N = flags.shape[1]
    em, _ = cell_tables(N, 2, 2)
    zero = (flags == 0).to(torch.float32)  # B x N
    return (zero @ em.t() == 0).to(torch.float32)  # B x E


def cell_flags(flags: Tensor, d_min: int, d_max: int) -> Tensor:
    """flags_right of get_rank2_flags in closed form.  cc_utils.py:527-557."""
    N = flags.shape[1]
    _, cm = cell_tables(N, d_min, d_max)
    zero = (flags == 0).to(torch.float32)
    return (zero @ cm.t() == 0).to(torch.float32)  # B x K


def mask_rank2(rank2: Tensor, N: int, d_min: int, d_max: int, flags: Optional[Tensor]) -> Tensor:
    """cc_utils.py:560-591 (3-D and 4-D)."""
    if flags is None:
        return rank2
    fl, fr = edge_flags(flags), cell_flags(flags, d_min, d_max)
    if rank2.dim() == 4:
        fl, fr = fl.unsqueeze(1), fr.unsqueeze(1)
    return fl.unsqueeze(-1) * rank2 * fr.unsqueeze(-2)


def mask_hodge_adjs(h: Tensor, flags: Optional[Tensor]) -> Tensor:
    """cc_utils.py:1615-1641."""
    if flags is None:
        return h
    fe = edge_flags(flags)
    if h.dim() == 4:
        fe = fe.unsqueeze(1)
    return h * fe.unsqueeze(-1) * fe.unsqueeze(-2)


def symmetrize_noise(z: Tensor) -> Tensor:
    """triu(1) + transpose, as in gen_noise(sym=True) / prior_sampling_sym.
    graph_utils.py:172-174, sde.py:448-449."""
    z = z.triu(1)
    return z + z.transpose(-1, -2)


def quantize(t: Tensor, thr: float = 0.5) -> Tensor:
    """graph_utils.py:181-192."""
    return torch.where(t < thr, torch.zeros_like(t), torch.ones_like(t))


def quantize_mol(adjs: Tensor) -> Tensor:
    """graph_utils.py:195-213 (thresholds .5/1.5/2.5 -> {0,1,2,3}, int64)."""
    a = adjs.detach().clone()
    out = torch.zeros_like(a)
    out[a >= 0.5] = 1
    out[a >= 1.5] = 2
    out[a >= 2.5] = 3
    return out.to(torch.int64)


def mol_onehot(x: Tensor, adj: Tensor) -> Tuple[Tensor, Tensor]:
    """Tensor post-processing of Sampler_mol.sample between the sampler and gen_mol.  sampler.py:814-825."""
    s = quantize_mol(adj) - 1
    s[s == -1] = 3  # 0, 1, 2, 3 (no, S, D, T) -> 3, 0, 1, 2
    a = Fnn.one_hot(s, num_classes=4).permute(0, 3, 1, 2)
    xi = torch.where(x > 0.5, 1, 0)
    xi = torch.concat([xi, 1 - xi.sum(dim=-1, keepdim=True)], dim=-1)
    return xi, a


# --------------------------------------------------------------------------------------
# tensor utilities
# --------------------------------------------------------------------------------------
def pow_tensor(x: Tensor, cnum: int) -> Tensor:
    """[A, A^2, ..., A^cnum] stacked on dim 1.  graph_utils.py:274-292."""
    xs, cur = [x], x
    for _ in range(cnum - 1):
        cur = torch.bmm(cur, x)
        xs.append(cur)
    return torch.stack(xs, dim=1)


def pow_tensor_cc(rank2: Tensor, cnum: int, use_hodge_mask: bool = True) -> Tensor:
    """[F, H F, H H F, ...] with H = (F F^T) * (1 - I).  cc_utils.py:945-979."""
    H = rank2 @ rank2.transpose(-1, -2)
    if use_hodge_mask:
        E = H.shape[-1]
        H = H * (1.0 - torch.eye(E)).unsqueeze(0)
    xs, cur = [rank2], rank2
    for _ in range(cnum - 1):
        cur = torch.bmm(H, cur)
        xs.append(cur)
    return torch.stack(xs, dim=1)


def adj_to_hodgedual(adj: Tensor) -> Tensor:
    """Upper-triangular entries -> diag_embed (E x E, purely diagonal).  cc_utils.py:1503-1538."""
    N = adj.shape[-1]
    r, c = torch.triu_indices(N, N, offset=1)
    return torch.diag_embed(adj[..., r, c])


def hodgedual_to_adj(h: Tensor) -> Tensor:
    """Reads only the diagonal, scatters symmetrically.  cc_utils.py:1541-1588."""
    E = h.shape[-1]
    N = int((1 + math.isqrt(1 + 8 * E)) // 2)
    r, c = torch.triu_indices(N, N, offset=1)
    d = h.diagonal(dim1=-2, dim2=-1)
    out = torch.zeros(*h.shape[:-2], N, N)
    out[..., r, c] = d
    out[..., c, r] = d
    return out


# --------------------------------------------------------------------------------------
# layers  (state_dict based)
# --------------------------------------------------------------------------------------
def _sub(sd: SD, prefix: str) -> SD:
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


def mlp(sd: SD, x: Tensor, act: Callable[[Tensor], Tensor] = Fnn.elu) -> Tensor:
    """MLP.forward, use_bn=False.  layers.py:246-275.  Keys: linear.* or linears.<i>.*"""
    if "linear.weight" in sd:
        return Fnn.linear(x, sd["linear.weight"], sd["linear.bias"])
    n = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("linears."))
    h = x
    for i in range(n - 1):
        h = act(Fnn.linear(h, sd[f"linears.{i}.weight"], sd[f"linears.{i}.bias"]))
    return Fnn.linear(h, sd[f"linears.{n-1}.weight"], sd[f"linears.{n-1}.bias"])


def dense_gcn(sd: SD, x: Tensor, adj: Tensor) -> Tensor:
    """DenseGCNConv.forward (add_loop=True, improved=False).  layers.py:115-158."""
    N = adj.shape[-1]
    adj = adj.clone()
    idx = torch.arange(N)
    adj[:, idx, idx] = 1.0
    out = x @ sd["weight"]
    d = adj.sum(dim=-1).clamp(min=1).pow(-0.5)
    adj = d.unsqueeze(-1) * adj * d.unsqueeze(-2)
    return adj @ out + sd["bias"]


def dense_hcn(sd: SD, hodge_adj: Tensor, rank2: Tensor) -> Tensor:
    """DenseHCNConv.forward (no self loops).  hodge_layers.py:163-199."""
    out = rank2 @ sd["weight"]
    d = hodge_adj.sum(dim=-1).clamp(min=1).pow(-0.5)
    h = d.unsqueeze(-1) * hodge_adj * d.unsqueeze(-2)
    return h @ out + sd["bias"]


def _tanh_attention(Q: Tensor, K: Tensor, heads: int, out_dim: int) -> Tensor:
    """heads split on the feature dim, tanh(QK^T/sqrt(out_dim)), mean over heads,
    symmetrise.  attention.py:111-130 / hodge_attention.py:108-127."""
    attn_dim = Q.shape[-1]
    ds = attn_dim // heads
    Q_ = torch.cat(Q.split(ds, 2), 0)
    K_ = torch.cat(K.split(ds, 2), 0)
    A = torch.tanh(Q_.bmm(K_.transpose(1, 2)) / math.sqrt(out_dim))
    A = A.view(-1, Q.shape[0], Q.shape[1], Q.shape[1]).mean(dim=0)
    return (A + A.transpose(-1, -2)) / 2


def attention(sd: SD, x: Tensor, adj: Tensor, heads: int) -> Tuple[Tensor, Tensor]:
    """Attention.forward.  attention.py:84-132; conv == "GCN": Q, K = DenseGCNConv; conv == "MLP": Q, K = 2-layer
    tanh MLPs of x alone (attention.py:170-180).  V is a DenseGCNConv in both."""
    if "gnn_q.weight" in sd:
        Q = dense_gcn(_sub(sd, "gnn_q"), x, adj)
        K = dense_gcn(_sub(sd, "gnn_k"), x, adj)
    else:
        Q = mlp(_sub(sd, "gnn_q"), x, act=torch.tanh)
        K = mlp(_sub(sd, "gnn_k"), x, act=torch.tanh)
    V = dense_gcn(_sub(sd, "gnn_v"), x, adj)
    out_dim = sd["gnn_v.weight"].shape[1]
    return V, _tanh_attention(Q, K, heads, out_dim)


def attention_layer(sd: SD, x: Tensor, adj: Tensor, flags: Optional[Tensor], heads: int) -> Tuple[Tensor, Tensor]:
    """AttentionLayer.forward.  attention.py:270-304."""
    c_in = adj.shape[1]
    masks, xs = [], []
    for k in range(c_in):
        v, a = attention(_sub(sd, f"attn.{k}"), x, adj[:, k], heads)
        masks.append(a.unsqueeze(-1))
        xs.append(v)
    x_out = torch.tanh(mask_x(mlp(_sub(sd, "multi_channel"), torch.cat(xs, dim=-1)), flags))
    mlp_in = torch.cat([torch.cat(masks, dim=-1), adj.permute(0, 2, 3, 1)], dim=-1)
    m = mlp(_sub(sd, "mlp"), mlp_in).permute(0, 3, 1, 2)
    m = m + m.transpose(-1, -2)
    return x_out, mask_adjs(m, flags)


def hodge_attention(sd: SD, hodge_adj: Tensor, rank2: Tensor, heads: int) -> Tuple[Tensor, Tensor]:
    """HodgeAttention.forward with conv == "HCN" (ccnn_v = Identity).  hodge_attention.py:80-129."""
    Q = dense_hcn(_sub(sd, "ccnn_q"), hodge_adj, rank2)
    K = dense_hcn(_sub(sd, "ccnn_k"), hodge_adj, rank2)
    V = torch.bmm(hodge_adj, rank2)
    out_dim = rank2.shape[-1]  # HodgeAttention(K, attn_dim, K): out_dim = number of cells
    return V, _tanh_attention(Q, K, heads, out_dim)


def hodge_adj_attention_layer(
    sd: SD, hodge_adj: Tensor, rank2: Tensor, flags: Optional[Tensor], heads: int, N: int, d_min: int, d_max: int
) -> Tuple[Tensor, Tensor]:
    """HodgeAdjAttentionLayer.forward.  hodge_attention.py:290-325."""
    c_in = hodge_adj.shape[1]
    vals, atts = [], []
    for k in range(c_in):
        v, a = hodge_attention(_sub(sd, f"attn.{k}"), hodge_adj[:, k], rank2, heads)
        vals.append(v.unsqueeze(-1))
        atts.append(a.unsqueeze(-1))
    h = mask_hodge_adjs(mlp(_sub(sd, "mlp_attention"), torch.cat(atts, dim=-1)).permute(0, 3, 1, 2), flags)
    h = torch.tanh(h)
    h = h + h.transpose(-1, -2)
    r = mlp(_sub(sd, "mlp_value"), torch.cat(vals, dim=-1)).squeeze(-1)
    return h, mask_rank2(r, N, d_min, d_max, flags)


def hodge_network_layer(sd: SD, rank2c: Tensor, N: int, d_min: int, d_max: int, flags: Optional[Tensor]) -> Tensor:
    """HodgeNetworkLayer.forward.  hodge_layers.py:70-92."""
    out = mlp(_sub(sd, "layer"), rank2c.permute(0, 2, 3, 1)).permute(0, 3, 1, 2)
    return mask_rank2(out, N, d_min, d_max, flags)


# --------------------------------------------------------------------------------------
# score networks
# --------------------------------------------------------------------------------------
def _count(sd: SD, prefix: str) -> int:
    idx = [int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")]
    return 1 + max(idx) if idx else 0


def score_network_x(sd: SD, hp: dict, x: Tensor, adj: Tensor, flags: Optional[Tensor]) -> Tensor:
    """ScoreNetworkX.forward_graph.  ScoreNetwork_X.py:102-133."""
    xs = [x]
    for k in range(hp["depth"]):
        x = torch.tanh(dense_gcn(_sub(sd, f"layers.{k}"), x, adj))
        xs.append(x)
    out = mlp(_sub(sd, "final"), torch.cat(xs, dim=-1))
    return mask_x(out, flags)


def score_network_x_gmh(sd: SD, hp: dict, x: Tensor, adj: Tensor, flags: Optional[Tensor]) -> Tensor:
    """ScoreNetworkX_GMH.forward_graph.  ScoreNetwork_X.py:280-314: AttentionLayers on (x, A^1..A^c), tanh of every
    layer's node output (on top of the layer's own tanh), concat, 3-layer MLP, mask."""
    adjc = pow_tensor(adj, hp["c_init"])
    xs = [x]
    for k in range(hp["depth"]):
        x, adjc = attention_layer(_sub(sd, f"layers.{k}"), x, adjc, flags, hp["num_heads"])
        x = torch.tanh(x)
        xs.append(x)
    out = mlp(_sub(sd, "final"), torch.cat(xs, dim=-1))
    return mask_x(out, flags)


def _adj_trunk(sd: SD, hp: dict, x: Tensor, adj: Tensor, flags: Optional[Tensor]) -> List[Tensor]:
    adjc = pow_tensor(adj, hp["c_init"])
    lst = [adjc]
    for k in range(hp["num_layers"]):
        x, adjc = attention_layer(_sub(sd, f"layers.{k}"), x, adjc, flags, hp["num_heads"])
        lst.append(adjc)
    return lst


def score_network_a(sd: SD, hp: dict, x: Tensor, adj: Tensor, flags: Optional[Tensor]) -> Tensor:
    """ScoreNetworkA.forward_graph.  ScoreNetwork_A.py:505-541."""
    adjs = torch.cat(_adj_trunk(sd, hp, x, adj, flags), dim=1).permute(0, 2, 3, 1)
    score = mlp(_sub(sd, "final"), adjs).squeeze(-1)
    N = adj.shape[-1]
    score = score * (1.0 - torch.eye(N)).unsqueeze(0)
    return mask_adjs(score, flags)


def score_network_a_cc(sd: SD, hp: dict, x: Tensor, adj: Tensor, rank2: Tensor, flags: Optional[Tensor]) -> Tensor:
    """ScoreNetworkA_CC.forward.  ScoreNetwork_A_CC.py:275-332."""
    N = adj.shape[-1]
    adj_list = _adj_trunk(sd, hp, x, adj, flags)
    hodge = adj_to_hodgedual(adj_list[0])
    hodge_list = [hodge]
    r2 = rank2
    for k in range(hp["num_layers_h"]):
        hodge, r2 = hodge_adj_attention_layer(
            _sub(sd, f"layers_hodge.{k}"), hodge, r2, flags, hp["num_heads_h"], N, hp["d_min"], hp["d_max"]
        )
        hodge_list.append(hodge)
    adjs = torch.cat(adj_list, dim=1).permute(0, 2, 3, 1)
    hadj = hodgedual_to_adj(torch.cat(hodge_list, dim=1)).permute(0, 2, 3, 1)
    score = mlp(_sub(sd, "final"), torch.cat([adjs, hadj], dim=-1)).squeeze(-1)
    score = score * (1.0 - torch.eye(N)).unsqueeze(0)
    return mask_adjs(score, flags)


def baseline_block(sd: SD, hodge_adj: Tensor, rank2: Tensor) -> Tuple[Tensor, Tensor]:
    """BaselineBlock.forward: row-wise 2-layer MLP (E -> hidden -> E, elu) on the Hodge adjacency, tanh;
    rank2_out = that @ rank2, hodge_out = its symmetrisation.  hodge_layers.py:261-284."""
    h = torch.tanh(mlp(_sub(sd, "mlp_layer"), hodge_adj))
    return torch.bmm(h, rank2), (h + h.transpose(-1, -2)) / 2


def hodge_baseline_layer(sd: SD, hodge_adj: Tensor, rank2: Tensor, flags: Optional[Tensor], N: int, d_min: int,
                         d_max: int) -> Tuple[Tensor, Tensor]:
    """HodgeBaselineLayer.forward.  hodge_layers.py:380-416."""
    r2s, hs = [], []
    for k in range(hodge_adj.shape[1]):
        r, h = baseline_block(_sub(sd, f"layers.{k}"), hodge_adj[:, k], rank2)
        r2s.append(r.unsqueeze(-1))
        hs.append(h.unsqueeze(-1))
    out = mask_hodge_adjs(mlp(_sub(sd, "mlp_hodge"), torch.cat(hs, dim=-1)).permute(0, 3, 1, 2), flags)
    out = torch.tanh(out)
    out = out + out.transpose(-1, -2)
    r2 = mask_rank2(mlp(_sub(sd, "mlp_rank2"), torch.cat(r2s, dim=-1)).squeeze(-1), N, d_min, d_max, flags)
    return out, r2


def score_network_a_base_cc(sd: SD, hp: dict, x: Tensor, adj: Tensor, rank2: Tensor, flags: Optional[Tensor]) -> Tensor:
    """ScoreNetworkA_Base_CC.forward.  ScoreNetwork_A_Base_CC.py:266-323."""
    N = adj.shape[-1]
    adj_list = _adj_trunk(sd, hp, x, adj, flags)
    hodge = adj_to_hodgedual(adj_list[0])
    hodge_list = [hodge]
    r2 = rank2
    for k in range(hp["num_layers_h"]):
        hodge, r2 = hodge_baseline_layer(_sub(sd, f"layers_hodge.{k}"), hodge, r2, flags, N, hp["d_min"], hp["d_max"])
        hodge_list.append(hodge)
    adjs = torch.cat(adj_list, dim=1).permute(0, 2, 3, 1)
    hadj = hodgedual_to_adj(torch.cat(hodge_list, dim=1)).permute(0, 2, 3, 1)
    score = mlp(_sub(sd, "final"), torch.cat([adjs, hadj], dim=-1)).squeeze(-1)
    score = score * (1.0 - torch.eye(N)).unsqueeze(0)
    return mask_adjs(score, flags)


def score_network_f(sd: SD, hp: dict, rank2: Tensor, flags: Optional[Tensor]) -> Tensor:
    """ScoreNetworkF.forward (ignores x and adj).  ScoreNetwork_F.py:175-217."""
    N, d_min, d_max = hp["max_node_num"], hp["d_min"], hp["d_max"]
    r2c = pow_tensor_cc(rank2, hp["cnum"], hp.get("use_hodge_mask", True))
    lst, cur = [r2c], r2c
    for k in range(hp["num_layers"]):
        cur = hodge_network_layer(_sub(sd, f"layers.{k}"), cur, N, d_min, d_max, flags)
        lst.append(cur)
    r2s = torch.cat(lst, dim=1).permute(0, 2, 3, 1)
    score = mlp(_sub(sd, "final"), r2s).squeeze(-1)
    return mask_rank2(score, N, d_min, d_max, flags)


class Model:
    """A (state_dict, hyper-parameter) pair that is callable like the reference Modules:
    graph nets ``m(x, adj, flags)``, CC nets ``m(x, adj, rank2, flags)``."""

    def __init__(self, kind: str, hp: dict, sd: SD, is_cc: Optional[bool] = None):
        self.kind, self.hp = kind, dict(hp)
        self.sd = {k[7:] if k.startswith("module.") else k: v.detach().to(torch.float32) for k, v in sd.items()}
        self.is_cc = bool(hp.get("is_cc", False)) if is_cc is None else is_cc

    def __call__(self, *args):
        if self.kind == "ScoreNetworkX":
            x, adj, flags = (args[0], args[1], args[-1])
            return score_network_x(self.sd, self.hp, x, adj, flags)
        if self.kind == "ScoreNetworkX_GMH":
            x, adj, flags = (args[0], args[1], args[-1])
            return score_network_x_gmh(self.sd, self.hp, x, adj, flags)
        if self.kind == "ScoreNetworkA":
            x, adj, flags = (args[0], args[1], args[-1])
            return score_network_a(self.sd, self.hp, x, adj, flags)
        if self.kind == "ScoreNetworkA_CC":
            x, adj, rank2, flags = args
            return score_network_a_cc(self.sd, self.hp, x, adj, rank2, flags)
        if self.kind == "ScoreNetworkA_Base_CC":
            x, adj, rank2, flags = args
            return score_network_a_base_cc(self.sd, self.hp, x, adj, rank2, flags)
        if self.kind == "ScoreNetworkF":
            x, adj, rank2, flags = args
            return score_network_f(self.sd, self.hp, rank2, flags)
        raise NotImplementedError(self.kind)


# --------------------------------------------------------------------------------------
# SDEs   (sde.py:345-786)
# --------------------------------------------------------------------------------------
class VPSDE:
    """sde.py:345-503."""

    def __init__(self, beta_min=0.1, beta_max=20.0, N=1000):
        self.beta_0, self.beta_1, self.N, self.T = beta_min, beta_max, N, 1
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        return -0.5 * beta_t[:, None, None] * x, torch.sqrt(beta_t)

    def marginal_std(self, t):
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.sqrt(1.0 - torch.exp(2.0 * lmc))

    def discretize(self, x, t):
        ts = (t * (self.N - 1) / self.T).long()
        beta, alpha = self.discrete_betas[ts], self.alphas[ts]
        return torch.sqrt(alpha)[:, None, None] * x - x, torch.sqrt(beta)

    def transition(self, x, t, dt):
        lmc = 0.25 * dt * (2 * self.beta_0 + (2 * t + dt) * (self.beta_1 - self.beta_0))
        return torch.exp(-lmc[:, None, None]) * x, torch.sqrt(1.0 - torch.exp(2.0 * lmc))


class VESDE:
    """sde.py:506-669."""

    def __init__(self, sigma_min=0.01, sigma_max=50.0, N=1000):
        self.sigma_min, self.sigma_max, self.N, self.T = sigma_min, sigma_max, N, 1
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    def sde(self, x, t):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        g = sigma * torch.sqrt(torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min))))
        return torch.zeros_like(x), g

    def marginal_std(self, t):
        return self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def discretize(self, x, t):
        ts = (t * (self.N - 1) / self.T).long()
        sigma = self.discrete_sigmas[ts]
        adjacent = torch.where(ts == 0, torch.zeros_like(t), self.discrete_sigmas[ts - 1])
        return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)

    def transition(self, x, t, dt):
        std = torch.square(self.sigma_min * (self.sigma_max / self.sigma_min) ** t) - torch.square(
            self.sigma_min * (self.sigma_max / self.sigma_min) ** (t + dt)
        )
        return x, torch.sqrt(std)


class subVPSDE:
    """sde.py:672-786 (Euler discretisation inherited from SDE.discretize :93-111; std WITHOUT sqrt)."""

    def __init__(self, beta_min=0.1, beta_max=20.0, N=1000):
        self.beta_0, self.beta_1, self.N, self.T = beta_min, beta_max, N, 1
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        discount = 1.0 - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
        return -0.5 * beta_t[:, None, None] * x, torch.sqrt(beta_t * discount)

    def marginal_std(self, t):
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return 1 - torch.exp(2.0 * lmc)

    def discretize(self, x, t):
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt))


def make_sde(kind: str, beta_min: float, beta_max: float, num_scales: int):
    """load_sde: VE is built with sigma_min=beta_min, sigma_max=beta_max.  loader.py:242-275."""
    if kind == "VP":
        return VPSDE(beta_min, beta_max, num_scales)
    if kind == "VE":
        return VESDE(beta_min, beta_max, num_scales)
    if kind == "subVP":
        return subVPSDE(beta_min, beta_max, num_scales)
    raise NotImplementedError(f"SDE class {kind} not yet supported.")


def score_fn(sde, model: Callable, *state_and_flags, t: Tensor) -> Tensor:
    """get_score_fn / get_score_fn_cc (continuous=True): VP/subVP -> -model/std, VE -> model.
    losses.py:18-104, 107-198."""
    out = model(*state_and_flags)
    if isinstance(sde, (VPSDE, subVPSDE)):
        return -out / sde.marginal_std(t)[:, None, None]
    if isinstance(sde, VESDE):
        return out
    raise NotImplementedError(f"SDE class {sde.__class__.__name__} not supported.")


# --------------------------------------------------------------------------------------
# noise sources
# --------------------------------------------------------------------------------------
class NoiseSource:
    """Produces the *raw* Gaussian draws in reference order.  ``recorded``
    (a list of tensors) replays an injected stream; otherwise torch.randn with a generator."""

    def __init__(self, seed: Optional[int] = 0, recorded: Optional[List[Tensor]] = None):
        # seed=None -> torch's global CPU generator (used to replay the reference's own stream)
        self.gen = None if seed is None else torch.Generator().manual_seed(seed)
        self.recorded = recorded
        self.pos = 0
        self.log: List[Tensor] = []

    def draw(self, shape: Sequence[int]) -> Tensor:
        if self.recorded is not None:
            z = self.recorded[self.pos]
            assert tuple(z.shape) == tuple(shape), (z.shape, shape)
        else:
            z = torch.randn(*shape, generator=self.gen)
        self.pos += 1
        self.log.append(z)
        return z


# --------------------------------------------------------------------------------------
# samplers
# --------------------------------------------------------------------------------------
class _Ctx:
    def __init__(self, is_cc, N, d_min, d_max, flags, noise: NoiseSource):
        self.is_cc, self.N, self.d_min, self.d_max, self.flags, self.noise = is_cc, N, d_min, d_max, flags, noise

    def gen(self, obj: str, like: Tensor) -> Tensor:
        """gen_noise / gen_noise_rank2: raw draw, then symmetrise (adj) and mask.
        graph_utils.py:158-178, cc_utils.py:594-615."""
        z = self.noise.draw(like.shape)
        if obj == "adj":
            return mask_adjs(symmetrize_noise(z), self.flags)
        if obj == "x":
            return mask_x(z, self.flags)
        return mask_rank2(z, self.N, self.d_min, self.d_max, self.flags)


def _norm_mean(t: Tensor) -> Tensor:
    return torch.norm(t.reshape(t.shape[0], -1), dim=-1).mean()


def _langevin(ctx, obj, sde, score, state, cur, vec_t, snr, seps, n_steps):
    """LangevinCorrector.update_fn_*.  solver.py:661-807.  ``score(cur)`` evaluates the score
    with ``cur`` substituted for this object in the (pre-corrector) state."""
    if isinstance(sde, (VPSDE, subVPSDE)):
        alpha = sde.alphas[(vec_t * (sde.N - 1) / sde.T).long()]
    else:
        alpha = torch.ones_like(vec_t)
    mean = cur
    for _ in range(n_steps):
        grad = score(cur)
        noise = ctx.gen(obj, cur)
        step = (snr * _norm_mean(noise) / _norm_mean(grad)) ** 2 * 2 * alpha
        mean = cur + step[:, None, None] * grad
        cur = mean + torch.sqrt(step * 2)[:, None, None] * noise * seps
    return cur, mean


def _predict(ctx, kind, obj, sde, score, cur, vec_t, probability_flow):
    """ReverseDiffusionPredictor / EulerMaruyamaPredictor update + RSDE.  solver.py:210-463,
    sde.py:180-235, 265-340."""
    pf = 0.5 if probability_flow else 1.0
    if kind == "Reverse":
        f, G = sde.discretize(cur, vec_t)
        s = score(cur)
        rev_f = f - G[:, None, None] ** 2 * s * pf
        rev_G = torch.zeros_like(G) if probability_flow else G
        z = ctx.gen(obj, cur)
        mean = cur - rev_f
        return mean + rev_G[:, None, None] * z, mean
    if kind == "Euler":
        dt = -1.0 / sde.N
        z = ctx.gen(obj, cur)
        drift, diffusion = sde.sde(cur, vec_t)
        s = score(cur)
        drift = drift - diffusion[:, None, None] ** 2 * s * pf
        diffusion = torch.zeros_like(diffusion) if probability_flow else diffusion
        mean = cur + drift * dt
        return mean + diffusion[:, None, None] * np.sqrt(-dt) * z, mean
    raise NotImplementedError(f"Predictor {kind} not yet supported. Select from [Reverse, Euler].")


def _prior(ctx, shapes, sdes):
    """prior_sampling / prior_sampling_sym then mask.  solver.py:963-968, 1111-1118."""
    x = mask_x(ctx.noise.draw(shapes[0]), ctx.flags)
    adj = mask_adjs(symmetrize_noise(ctx.noise.draw(shapes[1])), ctx.flags)
    if ctx.is_cc:
        r2 = mask_rank2(ctx.noise.draw(shapes[2]), ctx.N, ctx.d_min, ctx.d_max, ctx.flags)
        return [x, adj, r2]
    return [x, adj]


def pc_sampler(
    models, sdes, shapes, flags, *, predictor="Euler", corrector="None", snr=0.1, scale_eps=1.0, n_steps=1,
    probability_flow=False, denoise=True, eps=1e-3, d_min=None, d_max=None, noise: Optional[NoiseSource] = None,
    max_steps: Optional[int] = None, record: Optional[list] = None,
):
    """get_pc_sampler(...)(models..., init_flags).  solver.py:856-1176.

    ``models``/``sdes``/``shapes`` are lists of length 2 (graph) or 3 (CC) in the order
    x, adj[, rank2].  ``max_steps`` truncates the loop (for bounded CPU timing / fixtures) on
    the REAL N-step schedule.  Returns (state_or_means, n_evals, per-step record).
    """
    if predictor not in ("Reverse", "Euler"):
        raise NotImplementedError(f"Predictor {predictor} not yet supported. Select from [Reverse, Euler].")
    if corrector not in ("Langevin", "None"):
        raise NotImplementedError(f"Corrector {corrector} not yet supported. Select from [Langevin, None].")
    is_cc = len(models) == 3
    noise = noise or NoiseSource(0)
    N = shapes[1][1]
    ctx = _Ctx(is_cc, N, d_min, d_max, flags, noise)
    names = ["x", "adj", "rank2"][: len(models)]
    with torch.no_grad():
        state = _prior(ctx, shapes, sdes)
        diff_steps = sdes[1].N
        timesteps = torch.linspace(sdes[1].T, eps, diff_steps)
        means = list(state)
        for i in range(diff_steps if max_steps is None else min(max_steps, diff_steps)):
            vec_t = torch.ones(shapes[1][0]) * timesteps[i]

            def make_score(k, base):
                def f(cur):
                    st = list(base)
                    st[k] = cur
                    return score_fn(sdes[k], models[k], *st, flags, t=vec_t)
                return f

            if corrector == "Langevin":
                base = list(state)
                outs = [
                    _langevin(ctx, names[k], sdes[k], make_score(k, base), base, base[k], vec_t, snr, scale_eps, n_steps)
                    for k in range(len(models))
                ]
                state = [o[0] for o in outs]
            base = list(state)
            outs = [
                _predict(ctx, predictor, names[k], sdes[k], make_score(k, base), base[k], vec_t, probability_flow)
                for k in range(len(models))
            ]
            state = [o[0] for o in outs]
            means = [o[1] for o in outs]
            if record is not None:
                record.append(([s.clone() for s in state], [m.clone() for m in means]))
    return (means if denoise else state), diff_steps * (n_steps + 1)


def s4_solver(
    models, sdes, shapes, flags, *, snr=0.1, scale_eps=1.0, denoise=True, eps=1e-3, d_min=None, d_max=None,
    noise: Optional[NoiseSource] = None, max_steps: Optional[int] = None, record: Optional[list] = None, **_unused,
):
    """S4_solver(...)(models..., init_flags).  solver.py:1179-1563."""
    is_cc = len(models) == 3
    noise = noise or NoiseSource(0)
    N = shapes[1][1]
    ctx = _Ctx(is_cc, N, d_min, d_max, flags, noise)
    names = ["x", "adj", "rank2"][: len(models)]
    nobj = len(models)
    with torch.no_grad():
        state = _prior(ctx, shapes, sdes)
        diff_steps = sdes[1].N
        timesteps = torch.linspace(sdes[1].T, eps, diff_steps)
        dt = -1.0 / diff_steps
        means = list(state)
        for i in range(diff_steps if max_steps is None else min(max_steps, diff_steps)):
            B = shapes[1][0]
            vec_t = torch.ones(B) * timesteps[i]
            vec_dt = torch.ones(B) * (dt / 2)
            scores = [score_fn(sdes[k], models[k], *state, flags, t=vec_t) for k in range(nobj)]
            sdrift = [-sdes[k].sde(state[k], vec_t)[1][:, None, None] ** 2 * scores[k] for k in range(nobj)]
            timestep = (vec_t * (sdes[0].N - 1) / sdes[0].T).long()
            # correction (solver.py:1299-1334 / 1448-1502): alpha only for VPSDE (not subVP)
            for k in range(nobj):
                z = ctx.gen(names[k], state[k])
                alpha = sdes[k].alphas[timestep] if isinstance(sdes[k], VPSDE) else torch.ones_like(vec_t)
                step = (snr * _norm_mean(z) / _norm_mean(scores[k])) ** 2 * 2 * alpha
                m = state[k] + step[:, None, None] * scores[k]
                state[k] = m + torch.sqrt(step * 2)[:, None, None] * z * scale_eps
            # prediction (solver.py:1336-1353 / 1504-1534)
            for k in range(nobj):
                mu, sig = sdes[k].transition(state[k], vec_t, vec_dt)
                state[k] = mu + sig[:, None, None] * ctx.gen(names[k], state[k])
            for k in range(nobj):
                state[k] = state[k] + sdrift[k] * dt
            for k in range(nobj):
                mu, sig = sdes[k].transition(state[k], vec_t + vec_dt, vec_dt)
                state[k] = mu + sig[:, None, None] * ctx.gen(names[k], state[k])
                means[k] = mu
            if record is not None:
                record.append(([s.clone() for s in state], [m.clone() for m in means]))
    return (means if denoise else state), 0
