"""Pack score-network weights (reference ``state_dict`` layout) into the contiguous
fp32 blob + topology descriptor the kernels read (include/ccsd_b200.h).

Accepted models: the reference's own ``torch.nn.Module`` instances (ScoreNetworkX, ScoreNetworkA,
ScoreNetworkA_CC, ScoreNetworkF -- possibly DataParallel-wrapped, ccsd/src/utils/loader.py:649-650),
this package's mirrors in ``ccsd_b200.models``, or any object with the same hyper-parameter
attributes and a ``state_dict()``.

Blob conventions (must match ccsd_b200/csrc/prims.cuh):
  * every tensor starts at a multiple of 4 floats;
  * Linear weights (out, in) are stored transposed as (in, out_pad), out_pad = round_up(out, 8),
    zero padded; biases have out_pad entries;
  * DenseGCNConv weights are already (in, out) (layers.py:87) -> (in, out_pad);
  * hodge q/k weights (K, attn_dim) (hodge_layers.py:135) are stored as projection ROWS of length
    K_pad = round_up(K, 4) so that rank2 @ W becomes extra columns of the Gram product.
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import _native as nat


def _unwrap(model):
    return model.module if hasattr(model, "module") and hasattr(model.module, "state_dict") else model


def _sd(model) -> Dict[str, np.ndarray]:
    sd = _unwrap(model).state_dict()
    out = {}
    for k, v in sd.items():
        if k.startswith("module."):  # loader.py:635-637
            k = k[7:]
        out[k] = v.detach().to("cpu", torch.float32).numpy()
    return out


def model_kind(model) -> str:
    m = _unwrap(model)
    return getattr(m, "model_type", None) or type(m).__name__


def _r8(v: int) -> int:
    return (v + 7) // 8 * 8


class Blob:
    def __init__(self):
        self.parts: List[np.ndarray] = []
        self.n = 0

    def add(self, a: np.ndarray) -> int:
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
        off = self.n
        pad = (-a.size) % 4
        self.parts.append(a)
        if pad:
            self.parts.append(np.zeros(pad, np.float32))
        self.n += a.size + pad
        return off

    def add_in_out(self, w_in_out: np.ndarray, bias: np.ndarray) -> Tuple[int, int]:
        din, dout = w_in_out.shape
        op = _r8(dout)
        wp = np.zeros((din, op), np.float32)
        wp[:, :dout] = w_in_out
        bp = np.zeros(op, np.float32)
        bp[:dout] = bias
        return self.add(wp), self.add(bp)

    def finish(self) -> np.ndarray:
        # tail padding so that 8-wide weight loads never run past the allocation
        self.parts.append(np.zeros(16, np.float32))
        self.n += 16
        return np.concatenate(self.parts) if self.parts else np.zeros(16, np.float32)


def _fill_mlp(dst: nat.Mlp, blob: Blob, sd: Dict[str, np.ndarray], prefix: str) -> None:
    """layers.py:205-218: ``linear`` when num_layers == 1 else ``linears.<i>``."""
    if f"{prefix}.linear.weight" in sd:
        ws = [(sd[f"{prefix}.linear.weight"], sd[f"{prefix}.linear.bias"])]
    else:
        ws, i = [], 0
        while f"{prefix}.linears.{i}.weight" in sd:
            ws.append((sd[f"{prefix}.linears.{i}.weight"], sd[f"{prefix}.linears.{i}.bias"]))
            i += 1
    if not ws:
        raise ValueError(f"no MLP weights under '{prefix}'")
    if len(ws) > nat.MAX_MLP:
        raise NotImplementedError(f"MLP '{prefix}' has {len(ws)} linears; at most {nat.MAX_MLP} are supported")
    if any(k.startswith(f"{prefix}.batch_norms") for k in sd):
        raise NotImplementedError("use_bn=True is not supported (no shipped config or checkpoint uses it)")
    dst.nl = len(ws)
    dst.din = ws[0][0].shape[1]
    dst.dout = ws[-1][0].shape[0]
    dst.dhid = ws[0][0].shape[0] if len(ws) > 1 else dst.dout
    for i, (w, b) in enumerate(ws):
        dst.w[i], dst.b[i] = blob.add_in_out(w.T, b)


def _fill_gcn(dst: nat.Gcn, blob: Blob, sd, prefix: str) -> None:
    w, b = sd[f"{prefix}.weight"], sd[f"{prefix}.bias"]
    dst.din, dst.dout = w.shape
    dst.w, dst.b = blob.add_in_out(w, b)


def _count(sd, prefix: str) -> int:
    idx = {int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")}
    return max(idx) + 1 if idx else 0


def _fill_attn_layer(ly: nat.AttnLayer, blob: Blob, sd, prefix: str, conv: str) -> int:
    """One AttentionLayer (attention.py:186-304) under `prefix`; returns c_in."""
    c_in = _count(sd, f"{prefix}.attn")
    if c_in > nat.MAX_CH:
        raise NotImplementedError(f"{c_in} channels per attention layer; at most {nat.MAX_CH} are supported")
    if conv not in ("GCN", "MLP"):
        raise NotImplementedError(f"Convolution layer {conv} not implemented.")   # attention.py:183
    ly.conv_mlp = 1 if conv == "MLP" else 0
    for c in range(c_in):
        if ly.conv_mlp:      # Q, K: MLP(2, in, 2 attn_dim, attn_dim, tanh) of x (attention.py:170-180)
            _fill_mlp(ly.qm[c], blob, sd, f"{prefix}.attn.{c}.gnn_q")
            _fill_mlp(ly.km[c], blob, sd, f"{prefix}.attn.{c}.gnn_k")
            if ly.qm[c].nl != 2 or ly.km[c].nl != 2:
                raise NotImplementedError("conv == 'MLP': the Q / K networks must be 2-layer MLPs")
        else:
            _fill_gcn(ly.q[c], blob, sd, f"{prefix}.attn.{c}.gnn_q")
            _fill_gcn(ly.k[c], blob, sd, f"{prefix}.attn.{c}.gnn_k")
        _fill_gcn(ly.v[c], blob, sd, f"{prefix}.attn.{c}.gnn_v")
    _fill_mlp(ly.mlp, blob, sd, f"{prefix}.mlp")
    _fill_mlp(ly.multi_channel, blob, sd, f"{prefix}.multi_channel")
    # value convolution folded with the channel's slice of multi_channel's first Linear (float64 on the host):
    # V_c W1_c = An x (W_v W1_c) + b_v W1_c  (attention.py:292; csrc/tc_attn.cuh)
    mcp = f"{prefix}.multi_channel"
    w1 = sd[f"{mcp}.linear.weight"] if f"{mcp}.linear.weight" in sd else sd[f"{mcp}.linears.0.weight"]   # (o1, c_in * nh)
    for c in range(c_in):
        wv = sd[f"{prefix}.attn.{c}.gnn_v.weight"].astype(np.float64)     # (kin, nh)
        bv = sd[f"{prefix}.attn.{c}.gnn_v.bias"].astype(np.float64)
        nh = wv.shape[1]
        w1c = w1[:, c * nh:(c + 1) * nh].astype(np.float64).T              # (nh, o1)
        ly.vw[c].din, ly.vw[c].dout = wv.shape[0], w1c.shape[1]
        ly.vw[c].w, ly.vw[c].b = blob.add_in_out((wv @ w1c).astype(np.float32), (bv @ w1c).astype(np.float32))
    ly.c_in, ly.c_out = c_in, ly.mlp.dout
    if ly.conv_mlp:
        ly.conv_in, ly.attn_dim = ly.qm[0].din, ly.qm[0].dout
    else:
        ly.conv_in, ly.attn_dim = ly.q[0].din, ly.q[0].dout
    ly.conv_out = ly.v[0].dout
    return c_in


def pack_netx(dst: nat.NetX, blob: Blob, model) -> None:
    """ScoreNetworkX (ScoreNetwork_X.py:22-133) or ScoreNetworkX_GMH (ScoreNetwork_X.py:156-341)."""
    m = _unwrap(model)
    sd = _sd(model)
    depth = _count(sd, "layers")
    if depth < 1 or depth > nat.MAX_LAYERS:
        raise NotImplementedError(f"ScoreNetworkX depth {depth} not in 1..{nat.MAX_LAYERS}")
    dst.depth = depth
    if model_kind(model) == "ScoreNetworkX_GMH":
        dst.gmh = 1
        dst.gmh_heads = int(getattr(m, "num_heads", 4))
        for k in range(depth):
            c_in = _fill_attn_layer(dst.glayer[k], blob, sd, f"layers.{k}", getattr(m, "conv", "GCN"))
            if k == 0:
                dst.gmh_c_init = c_in
        dst.nfeat = dst.glayer[0].conv_in
        dst.nhid = dst.glayer[0].conv_out
    else:
        for k in range(depth):
            _fill_gcn(dst.gcn[k], blob, sd, f"layers.{k}")
        dst.nfeat = dst.gcn[0].din
        dst.nhid = dst.gcn[0].dout
    dst.fdim = dst.nfeat + depth * dst.nhid
    _fill_mlp(dst.fin, blob, sd, "final")


def pack_neta(dst: nat.NetA, blob: Blob, model, K: int) -> None:
    """ScoreNetworkA (ScoreNetwork_A.py:370-541) or ScoreNetworkA_CC (ScoreNetwork_A_CC.py:24-332)."""
    m = _unwrap(model)
    sd = _sd(model)
    kind = model_kind(model)
    L = _count(sd, "layers")
    if L < 1 or L > nat.MAX_LAYERS:
        raise NotImplementedError(f"ScoreNetworkA num_layers {L} not in 1..{nat.MAX_LAYERS}")
    dst.num_layers = L
    dst.num_heads = int(m.num_heads)
    dst.is_cc = 1 if kind in ("ScoreNetworkA_CC", "ScoreNetworkA_Base_CC") else 0
    dst.base_cc = 1 if kind == "ScoreNetworkA_Base_CC" else 0
    fdim = 0
    for l in range(L):
        ly = dst.layer[l]
        c_in = _fill_attn_layer(ly, blob, sd, f"layers.{l}", getattr(m, "conv", "GCN"))
        if l == 0:
            dst.c_init = c_in
            fdim += c_in
        fdim += ly.c_out
    if dst.base_cc:
        # ScoreNetworkA_Base_CC: HodgeBaselineLayers (hodge_layers.py:287-416).  Their rank-2 outputs never reach the
        # adjacency score (ScoreNetwork_A_Base_CC.py:296-312 reads only the hodge adjacencies), so mlp_rank2 is not packed.
        Lh = _count(sd, "layers_hodge")
        if Lh < 1 or Lh > nat.MAX_HODGE:
            raise NotImplementedError(f"ScoreNetworkA_Base_CC num_layers_h = {Lh}: only 1 or 2 hodge layers are implemented")
        dst.num_layers_h = Lh
        for l in range(Lh):
            h = dst.hbase[l]
            c_in = _count(sd, f"layers_hodge.{l}.layers")
            if c_in > nat.MAX_CH:
                raise NotImplementedError("more than 8 hodge channels")
            for c in range(c_in):
                pre = f"layers_hodge.{l}.layers.{c}.mlp_layer.linears"
                if f"{pre}.2.weight" in sd:
                    raise NotImplementedError("BaselineBlock MLP with more than 2 Linears")
                w1, b1 = sd[f"{pre}.0.weight"], sd[f"{pre}.0.bias"]       # (hid, E), (hid)
                w2, b2 = sd[f"{pre}.1.weight"], sd[f"{pre}.1.bias"]       # (E, hid), (E)
                hid = w1.shape[0]
                if hid > 32:
                    raise NotImplementedError("BaselineBlock hidden width > 32")
                h.w1[c], h.b1[c] = blob.add_in_out(w1.T, b1)
                w2p = np.zeros((w2.shape[0], _r8(hid)), np.float32)
                w2p[:, :hid] = w2
                h.w2[c], h.b2[c] = blob.add(w2p), blob.add(b2)
                h.hid = hid
            _fill_mlp(h.mlp_hodge, blob, sd, f"layers_hodge.{l}.mlp_hodge")
            h.c_in, h.c_out = c_in, h.mlp_hodge.dout
            fdim += h.c_out
        fdim += dst.c_init
    elif dst.is_cc:
        if getattr(m, "conv_hodge", "HCN") != "HCN":
            raise NotImplementedError("conv_hodge == 'MLP' is not implemented (no shipped config uses it)")
        Lh = _count(sd, "layers_hodge")
        if Lh < 1 or Lh > nat.MAX_HODGE:
            raise NotImplementedError(
                f"ScoreNetworkA_CC num_layers_h = {Lh}: only 1 or 2 hodge layers are implemented"
            )
        dst.num_layers_h = Lh
        dst.num_heads_h = int(m.num_heads_h)
        Kw = (K + 3) // 4 * 4
        rows = []
        for l in range(Lh):
            h = dst.hodge[l]
            c_in = _count(sd, f"layers_hodge.{l}.attn")
            if c_in > nat.MAX_CH:
                raise NotImplementedError("more than 8 hodge channels")
            ad = None
            for c in range(c_in):
                wq = sd[f"layers_hodge.{l}.attn.{c}.ccnn_q.weight"]  # (K, ad)
                wk = sd[f"layers_hodge.{l}.attn.{c}.ccnn_k.weight"]
                if wq.shape[0] != K:
                    raise ValueError(f"hodge weight rows {wq.shape[0]} != K {K}")
                ad = wq.shape[1]
                for w in (wq, wk):
                    r = np.zeros((ad, Kw), np.float32)
                    r[:, :K] = w.T
                    rows.append(r)
                h.bq[c] = blob.add(np.pad(sd[f"layers_hodge.{l}.attn.{c}.ccnn_q.bias"], (0, 8)))
                h.bk[c] = blob.add(np.pad(sd[f"layers_hodge.{l}.attn.{c}.ccnn_k.bias"], (0, 8)))
            _fill_mlp(h.mlp_attention, blob, sd, f"layers_hodge.{l}.mlp_attention")
            _fill_mlp(h.mlp_value, blob, sd, f"layers_hodge.{l}.mlp_value")
            h.c_in, h.c_out, h.attn_dim, h.proj_row = c_in, h.mlp_attention.dout, ad, 0
            dst.n_proj_rows[l] = c_in * 2 * ad
            fdim += h.c_out
        fdim += dst.c_init
        dst.proj_w = blob.add(np.concatenate(rows, axis=0))
    _fill_mlp(dst.fin, blob, sd, "final")
    dst.fdim = dst.fin.din
    if fdim != dst.fdim:
        raise ValueError(
            f"ScoreNetworkA: final MLP expects {dst.fdim} channels but the layer stack produces {fdim} "
            "(the reference would fail with a shape error on this configuration)"
        )


def pack_netf(dst: nat.NetF, blob: Blob, model) -> None:
    """ScoreNetworkF (ScoreNetwork_F.py:24-217)."""
    m = _unwrap(model)
    sd = _sd(model)
    L = _count(sd, "layers")
    if L < 1 or L > nat.MAX_F_LAYERS:
        raise NotImplementedError(f"ScoreNetworkF num_layers {L} not in 1..{nat.MAX_F_LAYERS}")
    dst.num_layers = L
    for l in range(L):
        _fill_mlp(dst.layer[l], blob, sd, f"layers.{l}.layer")
    _fill_mlp(dst.fin, blob, sd, "final")
    dst.cnum = dst.layer[0].din
    dst.fdim = dst.fin.din
    dst.use_hodge_mask = 1 if getattr(m, "use_hodge_mask", True) else 0
    if int(getattr(m, "cnum", dst.cnum)) != dst.cnum:
        raise ValueError("ScoreNetworkF: cnum attribute disagrees with the first layer's input width")
    # affine fold: all MLPs single Linears -> score = m * (a f + b (H f) + c)   (masks are {0,1})
    dst.affine = 0
    if dst.cnum == 2 and dst.fin.nl == 1 and all(dst.layer[l].nl == 1 for l in range(L)):
        def lin(prefix):
            return sd[f"{prefix}.linear.weight"].astype(np.float64), sd[f"{prefix}.linear.bias"].astype(np.float64)
        A = np.eye(2)            # current channels as an affine map of v = [f, H f]:  cur = A v + c
        c = np.zeros(2)
        rows_A, rows_c = [A], [c]
        for l in range(L):
            W, b = lin(f"layers.{l}.layer")
            A, c = W @ A, W @ c + b
            rows_A.append(A)
            rows_c.append(c)
        Wf, bf = lin("final")
        A_all, c_all = np.concatenate(rows_A, 0), np.concatenate(rows_c, 0)
        coef = Wf @ A_all            # (1, 2)
        gamma = float((Wf @ c_all + bf).reshape(-1)[0])
        dst.affine = 1
        dst.aff[0], dst.aff[1], dst.aff[2] = float(coef[0, 0]), float(coef[0, 1]), gamma


def rank2_dim(N: int, d_min: int, d_max: int) -> Tuple[int, int]:
    """cc_utils.py:268-283."""
    return (N * (N - 1)) // 2, sum(math.comb(N, d) for d in range(d_min, d_max + 1))
