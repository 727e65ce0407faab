"""Score-network modules with the reference's constructor signatures, attribute names and
``state_dict`` layout (ccsd/src/models/ScoreNetwork_X.py:22, ScoreNetwork_A.py:370,
ScoreNetwork_A_CC.py:24, ScoreNetwork_F.py:24; layers.py:57,161; attention.py:21,186;
hodge_layers.py:17,114; hodge_attention.py:18,185), so that ``load_model(params)`` +
``load_state_dict(checkpoint)`` works without the reference package (ccsd/src/utils/loader.py:70-100,
619-653).

They are parameter containers: ``forward`` evaluates the network with the CUDA kernels through the
C ABI's score seam (``ccsd_score_eval``) -- there is no PyTorch/CPU forward.  The samplers never call
``forward``; they pack the weights once per plan (ccsd_b200/packer.py).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
from torch import nn


def _glorot(t: torch.Tensor) -> None:
    """layers.py:20-29."""
    stdv = math.sqrt(6.0 / (t.size(-2) + t.size(-1)))
    t.data.uniform_(-stdv, stdv)


class DenseGCNConv(nn.Module):
    """Parameters of layers.py:57-103: weight (in, out), bias (out)."""

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        _glorot(self.weight)


DenseHCNConv = DenseGCNConv  # hodge_layers.py:114-151: same parameter layout


class MLP(nn.Module):
    """Parameters of layers.py:161-244 (use_bn=False): ``linear`` if num_layers == 1 else ``linears``."""

    def __init__(self, num_layers: int, input_dim: int, hidden_dim: int, output_dim: int, use_bn: bool = False) -> None:
        super().__init__()
        if use_bn:
            raise NotImplementedError("use_bn=True is not supported (no shipped config or checkpoint uses it)")
        if num_layers < 1:
            raise ValueError("Number of layers should be greater of equal to 1.")
        self.num_layers, self.input_dim, self.hidden_dim, self.output_dim = num_layers, input_dim, hidden_dim, output_dim
        if num_layers == 1:
            self.linear = nn.Linear(input_dim, output_dim)
            _glorot(self.linear.weight)
            self.linear.bias.data.zero_()
        else:
            dims = [input_dim] + [hidden_dim] * (num_layers - 1) + [output_dim]
            self.linears = nn.ModuleList(nn.Linear(a, b) for a, b in zip(dims[:-1], dims[1:]))
            for layer in range(num_layers - 1):  # the last Linear keeps torch's default init (layers.py:238-241)
                _glorot(self.linears[layer].weight)
                self.linears[layer].bias.data.zero_()


class Attention(nn.Module):
    def __init__(self, in_dim: int, attn_dim: int, out_dim: int, num_heads: int = 4, conv: str = "GCN") -> None:
        super().__init__()
        self.num_heads, self.attn_dim, self.out_dim, self.conv = num_heads, attn_dim, out_dim, conv
        if conv == "GCN":       # attention.py:164-168
            self.gnn_q, self.gnn_k = DenseGCNConv(in_dim, attn_dim), DenseGCNConv(in_dim, attn_dim)
        elif conv == "MLP":     # attention.py:170-180
            self.gnn_q, self.gnn_k = MLP(2, in_dim, 2 * attn_dim, attn_dim), MLP(2, in_dim, 2 * attn_dim, attn_dim)
        else:
            raise NotImplementedError(f"Convolution layer {conv} not implemented.")
        self.gnn_v = DenseGCNConv(in_dim, out_dim)


class AttentionLayer(nn.Module):
    """attention.py:186-248."""

    def __init__(self, num_linears, conv_input_dim, attn_dim, conv_output_dim, input_dim, output_dim, num_heads=4,
                 conv="GCN", use_bn=False) -> None:
        super().__init__()
        self.attn = nn.ModuleList(Attention(conv_input_dim, attn_dim, conv_output_dim, num_heads, conv) for _ in range(input_dim))
        self.hidden_dim = 2 * max(input_dim, output_dim)
        self.mlp = MLP(num_linears, 2 * input_dim, self.hidden_dim, output_dim, use_bn)
        self.multi_channel = MLP(2, input_dim * conv_output_dim, self.hidden_dim, conv_output_dim, use_bn)


class HodgeAttention(nn.Module):
    def __init__(self, in_dim: int, attn_dim: int, out_dim: int, num_heads: int = 4, conv: str = "HCN") -> None:
        super().__init__()
        if conv != "HCN":
            raise NotImplementedError(f"Convolution layer {conv} not implemented.")
        self.ccnn_q, self.ccnn_k, self.ccnn_v = DenseHCNConv(in_dim, attn_dim), DenseHCNConv(in_dim, attn_dim), nn.Identity()


class HodgeAdjAttentionLayer(nn.Module):
    """hodge_attention.py:185-260."""

    def __init__(self, num_linears, input_dim, attn_dim, conv_output_dim, N, d_min, d_max, num_heads=4, conv="HCN",
                 use_bn=False) -> None:
        super().__init__()
        from .packer import rank2_dim

        self.K = rank2_dim(N, d_min, d_max)[1]
        self.attn = nn.ModuleList(HodgeAttention(self.K, attn_dim, self.K, num_heads, conv) for _ in range(input_dim))
        self.hidden_dim = 2 * max(input_dim, conv_output_dim)
        self.mlp_value = MLP(num_linears, input_dim, self.hidden_dim, 1, use_bn)
        self.mlp_attention = MLP(num_linears, input_dim, self.hidden_dim, conv_output_dim, use_bn)


class BaselineBlock(nn.Module):
    """hodge_layers.py:202-245."""

    def __init__(self, in_dim: int, hidden_dim: int, out_dim: int) -> None:
        super().__init__()
        self.mlp_layer = MLP(2, in_dim, hidden_dim, out_dim, False)


class HodgeBaselineLayer(nn.Module):
    """hodge_layers.py:287-357."""

    def __init__(self, num_linears, input_dim, hidden_dim, conv_output_dim, N, d_min, d_max, use_bn=False) -> None:
        super().__init__()
        from .packer import rank2_dim

        E = rank2_dim(N, d_min, d_max)[0]
        self.layers = nn.ModuleList(BaselineBlock(E, hidden_dim, E) for _ in range(input_dim))
        self.hidden_dim_mlp = 2 * max(input_dim, conv_output_dim)
        self.mlp_rank2 = MLP(num_linears, input_dim, self.hidden_dim_mlp, 1, use_bn)
        self.mlp_hodge = MLP(num_linears, input_dim, self.hidden_dim_mlp, conv_output_dim, use_bn)


class HodgeNetworkLayer(nn.Module):
    """hodge_layers.py:17-63."""

    def __init__(self, num_linears, input_dim, nhid, output_dim, d_min, d_max, use_bn=False) -> None:
        super().__init__()
        self.layer = MLP(num_linears, input_dim, nhid, output_dim, use_bn)


# ---------------------------------------------------------------------------------------------
class _ScoreNet(nn.Module):
    """Shared forward: evaluate this one network on the GPU through the C ABI score seam."""

    _which = 0

    def _engine(self, B: int, x: torch.Tensor, rank2: Optional[torch.Tensor]):
        from .sde import VPSDE
        from .solver import Engine

        key = (B, x.device)
        cache: Dict[Tuple, object] = self.__dict__.setdefault("_engines", {})
        models = [None, None, None]
        models[self._which] = self
        eng = cache.get(key)
        if eng is not None and not eng.refresh_weights(models[: len(eng.shapes)]):   # live parameters, every call
            eng = None
        if eng is None:
            shapes = [(B, x.shape[1], x.shape[2]), (B, x.shape[1], x.shape[1])]
            kw = {}
            if rank2 is not None:
                shapes.append(tuple(rank2.shape))
                kw = dict(d_min=self.d_min, d_max=self.d_max)
            sde = VPSDE(0.1, 1.0, 1000)
            cache.clear()
            eng = cache[key] = Engine(models[: len(shapes)], [sde] * len(shapes), shapes, sampler="PC", device=x.device, **kw)
        return eng

    def _score(self, x, adj, rank2, flags):
        from . import _native as nat
        if not x.is_cuda and not nat.is_emulation():   # (the host-emulation build is the test-suite's, see _native.py)
            raise RuntimeError("ccsd_b200 score networks evaluate on a CUDA device only (no CPU fallback)")
        eng = self._engine(x.shape[0], x, rank2)
        return eng.score(self._which, x, adj, rank2, flags)


class ScoreNetworkX(_ScoreNet):
    """ScoreNetwork_X.py:22-153."""

    _which = 0

    def __init__(self, max_feat_num: int, depth: int, nhid: int, use_bn: bool = False, is_cc: bool = False) -> None:
        super().__init__()
        self.nfeat, self.depth, self.nhid, self.use_bn, self.is_cc = max_feat_num, depth, nhid, use_bn, is_cc
        self.layers = nn.ModuleList(DenseGCNConv(max_feat_num if k == 0 else nhid, nhid) for k in range(depth))
        self.fdim = max_feat_num + depth * nhid
        self.final = MLP(3, self.fdim, 2 * self.fdim, max_feat_num, use_bn)

    def forward(self, x, adj, *rest):
        flags = rest[-1] if rest else None  # forward_cc(x, adj, rank2, flags) ignores rank2 (ScoreNetwork_X.py:135-153)
        return self._score(x, adj, None, flags)


class ScoreNetworkX_GMH(_ScoreNet):
    """ScoreNetwork_X.py:156-341."""

    _which = 0

    def __init__(self, max_feat_num, depth, nhid, num_linears, c_init, c_hid, c_final, adim, num_heads=4, conv="GCN",
                 use_bn=False, is_cc=False) -> None:
        super().__init__()
        self.nfeat, self.depth, self.nhid, self.num_linears = max_feat_num, depth, nhid, num_linears
        self.c_init, self.c_hid, self.c_final, self.adim = c_init, c_hid, c_final, adim
        self.num_heads, self.conv, self.use_bn, self.is_cc = num_heads, conv, use_bn, is_cc
        ls = []
        for k in range(depth):
            if k == 0:
                ls.append(AttentionLayer(num_linears, max_feat_num, nhid, nhid, c_init, c_hid, num_heads, conv, use_bn))
            elif k == depth - 1:
                ls.append(AttentionLayer(num_linears, nhid, adim, nhid, c_hid, c_final, num_heads, conv, use_bn))
            else:
                ls.append(AttentionLayer(num_linears, nhid, adim, nhid, c_hid, c_hid, num_heads, conv, use_bn))
        self.layers = nn.ModuleList(ls)
        self.fdim = max_feat_num + depth * nhid
        self.final = MLP(3, self.fdim, 2 * self.fdim, max_feat_num, use_bn)

    def forward(self, x, adj, *rest):
        flags = rest[-1] if rest else None
        return self._score(x, adj, None, flags)


class ScoreNetworkA(_ScoreNet):
    """ScoreNetwork_A.py:370-561."""

    _which = 1

    def __init__(self, max_feat_num, max_node_num, nhid, num_layers, num_linears, c_init, c_hid, c_final, adim,
                 num_heads=4, conv="GCN", use_bn=False, is_cc=False) -> None:
        super().__init__()
        self.max_feat_num, self.max_node_num, self.nhid = max_feat_num, max_node_num, nhid
        self.num_layers, self.num_linears = num_layers, num_linears
        self.c_init, self.c_hid, self.c_final, self.adim = c_init, c_hid, c_final, adim
        self.num_heads, self.conv, self.use_bn, self.is_cc = num_heads, conv, use_bn, is_cc
        self.layers = nn.ModuleList(self._trunk())
        self.fdim = c_hid * (num_layers - 1) + c_final + c_init
        self.final = MLP(3, self.fdim, 2 * self.fdim, 1, use_bn)

    def _trunk(self):
        L = self.num_layers
        out = []
        for k in range(L):
            if k == 0:
                out.append(AttentionLayer(self.num_linears, self.max_feat_num, self.nhid, self.nhid, self.c_init, self.c_hid,
                                          self.num_heads, self.conv, self.use_bn))
            elif k == L - 1:
                out.append(AttentionLayer(self.num_linears, self.nhid, self.adim, self.nhid, self.c_hid, self.c_final,
                                          self.num_heads, self.conv, self.use_bn))
            else:
                out.append(AttentionLayer(self.num_linears, self.nhid, self.adim, self.nhid, self.c_hid, self.c_hid,
                                          self.num_heads, self.conv, self.use_bn))
        return out

    def forward(self, x, adj, *rest):
        flags = rest[-1] if rest else None
        return self._score(x, adj, None, flags)


class ScoreNetworkA_CC(ScoreNetworkA):
    """ScoreNetwork_A_CC.py:24-332."""

    def __init__(self, max_feat_num, max_node_num, d_min, d_max, nhid, nhid_h, num_layers, num_layers_h, num_linears,
                 num_linears_h, c_init, c_hid, c_hid_h, c_final, c_final_h, adim, adim_h, num_heads=4, num_heads_h=4,
                 conv="GCN", conv_hodge="HCN", use_bn=False, is_cc=True) -> None:
        if not is_cc:
            raise ValueError("ScoreNetworkA_CC is only for combinatorial complexes")
        nn.Module.__init__(self)
        self.max_feat_num, self.max_node_num, self.N, self.d_min, self.d_max = max_feat_num, max_node_num, max_node_num, d_min, d_max
        self.nhid, self.nhid_h, self.num_layers, self.num_layers_h = nhid, nhid_h, num_layers, num_layers_h
        self.num_linears, self.num_linears_h = num_linears, num_linears_h
        self.c_init, self.c_hid, self.c_hid_h, self.c_final, self.c_final_h = c_init, c_hid, c_hid_h, c_final, c_final_h
        self.adim, self.adim_h, self.num_heads, self.num_heads_h = adim, adim_h, num_heads, num_heads_h
        self.conv, self.conv_hodge, self.use_bn, self.is_cc = conv, conv_hodge, use_bn, is_cc
        self.layers = nn.ModuleList(self._trunk())
        hl = []
        for k in range(num_layers_h):
            if k == 0:
                hl.append(HodgeAdjAttentionLayer(num_linears_h, c_init, nhid_h, c_hid_h, max_node_num, d_min, d_max,
                                                 num_heads_h, conv_hodge, use_bn))
            elif k == num_layers_h - 1:
                hl.append(HodgeAdjAttentionLayer(num_linears_h, c_hid_h, adim_h, c_final_h, max_node_num, d_min, d_max,
                                                 num_heads_h, conv_hodge, use_bn))
            else:
                hl.append(HodgeAdjAttentionLayer(num_linears_h, c_hid_h, adim_h, c_hid_h, max_node_num, d_min, d_max,
                                                 num_heads_h, conv_hodge, use_bn))
        self.layers_hodge = nn.ModuleList(hl)
        self.fdim = c_hid * (num_layers - 1) + c_final + c_init + c_hid_h * (num_layers_h - 1) + c_final_h + c_init
        self.final = MLP(3, self.fdim, 2 * self.fdim, 1, use_bn)

    def forward(self, x, adj, rank2, flags=None):
        return self._score(x, adj, rank2, flags)


class ScoreNetworkA_Base_CC(ScoreNetworkA):
    """ScoreNetwork_A_Base_CC.py:24-323."""

    def __init__(self, max_feat_num, max_node_num, d_min, d_max, nhid, nhid_h, num_layers, num_layers_h, num_linears,
                 num_linears_h, c_init, c_hid, c_hid_h, c_final, c_final_h, adim, hidden_h, num_heads=4, conv="GCN",
                 use_bn=False, is_cc=True) -> None:
        if not is_cc:
            raise ValueError("ScoreNetworkA_Base_CC is only for combinatorial complexes")
        nn.Module.__init__(self)
        self.max_feat_num, self.max_node_num, self.N, self.d_min, self.d_max = max_feat_num, max_node_num, max_node_num, d_min, d_max
        self.nhid, self.nhid_h, self.num_layers, self.num_layers_h = nhid, nhid_h, num_layers, num_layers_h
        self.num_linears, self.num_linears_h = num_linears, num_linears_h
        self.c_init, self.c_hid, self.c_hid_h, self.c_final, self.c_final_h = c_init, c_hid, c_hid_h, c_final, c_final_h
        self.adim, self.hidden_h, self.num_heads = adim, hidden_h, num_heads
        self.conv, self.use_bn, self.is_cc = conv, use_bn, is_cc
        self.layers = nn.ModuleList(self._trunk())
        hl = []
        for k in range(num_layers_h):
            if k == 0:
                hl.append(HodgeBaselineLayer(num_linears_h, c_init, nhid_h, c_hid_h, max_node_num, d_min, d_max, use_bn))
            elif k == num_layers_h - 1:
                hl.append(HodgeBaselineLayer(num_linears_h, c_hid_h, hidden_h, c_final_h, max_node_num, d_min, d_max, use_bn))
            else:
                hl.append(HodgeBaselineLayer(num_linears_h, c_hid_h, hidden_h, c_hid_h, max_node_num, d_min, d_max, use_bn))
        self.layers_hodge = nn.ModuleList(hl)
        self.fdim = c_hid * (num_layers - 1) + c_final + c_init + c_hid_h * (num_layers_h - 1) + c_final_h + c_init
        self.final = MLP(3, self.fdim, 2 * self.fdim, 1, use_bn)

    def forward(self, x, adj, rank2, flags=None):
        return self._score(x, adj, rank2, flags)


class ScoreNetworkF(_ScoreNet):
    """ScoreNetwork_F.py:24-217."""

    _which = 2

    def __init__(self, num_layers_mlp, num_layers, num_linears, nhid, c_hid, c_final, cnum, max_node_num, d_min, d_max,
                 use_hodge_mask=True, use_bn=False, is_cc=True) -> None:
        super().__init__()
        self.num_layers_mlp, self.num_layers, self.num_linears, self.nhid = num_layers_mlp, num_layers, num_linears, nhid
        self.c_hid, self.c_final, self.cnum, self.max_node_num = c_hid, c_final, cnum, max_node_num
        self.d_min, self.d_max, self.use_hodge_mask, self.use_bn, self.is_cc = d_min, d_max, use_hodge_mask, use_bn, is_cc
        ls = []
        for k in range(num_layers):
            cin = cnum if k == 0 else c_hid
            cout = c_hid if (k == 0 or k < num_layers - 1) else c_final
            ls.append(HodgeNetworkLayer(num_linears, cin, nhid, cout, d_min, d_max, use_bn))
        self.layers = nn.ModuleList(ls)
        self.fdim = c_hid * (num_layers - 1) + c_final + cnum
        self.final = MLP(num_layers_mlp, self.fdim, 2 * self.fdim, 1, use_bn)

    def forward(self, x, adj, rank2, flags=None):
        return self._score(x, adj, rank2, flags)


def load_model(params: dict) -> nn.Module:
    """ccsd/src/utils/loader.py:70-100."""
    p = dict(params)
    model_type = p.pop("model_type", None)
    table = {"ScoreNetworkX": ScoreNetworkX, "ScoreNetworkX_GMH": ScoreNetworkX_GMH, "ScoreNetworkA": ScoreNetworkA, "ScoreNetworkA_CC": ScoreNetworkA_CC,
             "ScoreNetworkA_Base_CC": ScoreNetworkA_Base_CC, "ScoreNetworkF": ScoreNetworkF}
    if model_type not in table:
        raise ValueError(
            f"Model Name <{model_type}> is unknown. Please select from [ScoreNetworkX, ScoreNetworkX_GMH, ScoreNetworkA, "
            "ScoreNetworkA_CC, ScoreNetworkA_Base_CC, ScoreNetworkF]")
    return table[model_type](**p)
