"""Drop-in ``TransformerEncoder`` / ``MultiHeadSelfAttention`` (reference: layers.py:15-103) on libvitb200 kernels.

Same constructor signatures, parameter names (``la1, attention.{Wq,Wk,Wv,out_project}, la2, mlp.0, mlp.3``),
``save_attn_map`` / ``get_attention_map()`` protocol and (B,T,F)->(B,T,F) forward as the reference; the sub-modules
only hold parameters — the arithmetic runs in hand-written sm_100a kernels through one ``autograd.Function`` per
block.  There is no PyTorch fallback: without a B200 and the built library, forward raises.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .params import FlatLayout, FlatStore, LayerViews, layer_entries

_PRECISION = os.environ.get("VITB_PRECISION", "bf16")


def set_precision(p: str) -> None:
    """'bf16' (default: bf16 storage + tensor cores, fp32 accumulation) or 'fp32' (check mode: fp32 SIMT kernels)."""
    global _PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def act_dtype() -> torch.dtype:
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32


def _new_drop_seed() -> int:
    """Seed of a module's dropout stream, drawn from torch's global CPU generator (so `torch.manual_seed` makes runs repeatable)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


class _DropState:
    """Mixin for modules that own an nn.Dropout(p) stream: a seed and a count of training-mode forward calls (the `step` of
    the counter-based mask generator, so every call draws fresh masks and backward can regenerate them)."""

    _drop_seed: Optional[int]
    _drop_calls: int

    def _next_drop(self, p: float, training: bool) -> Optional[Fn.Drop]:
        if not training or p <= 0.0:
            return None
        if not 0.0 < p < 1.0:
            raise ValueError(f"dropout probability has to be in [0, 1), got {p}")
        if self._drop_seed is None:
            self._drop_seed = _new_drop_seed()
        self._drop_calls += 1  # the first training call is step 1, like TrainEngine's step count
        return Fn.Drop(p=float(p), seed=self._drop_seed, step=self._drop_calls)


class _FlatRoot:
    """Mixin: a module that can own the flat storage of its parameter tree."""

    _store: Optional[FlatStore]

    def _layout(self) -> FlatLayout:  # pragma: no cover - abstract
        raise NotImplementedError

    def _after_pack(self, store: FlatStore) -> None:
        pass

    def _ensure_packed(self) -> FlatStore:
        st = getattr(self, "_store", None)
        if st is None or not st.consistent():
            ops.require_device()
            p0 = next(self.parameters())
            if not p0.is_cuda:
                raise RuntimeError("vit-cifar_b200 modules run on CUDA only: call .cuda() first (no CPU fallback)")
            st = FlatStore(self, self._layout())
            object.__setattr__(self, "_store", st)
            self._after_pack(st)
        return st

    def _compute_buffer(self, st: FlatStore, refresh: bool = True) -> torch.Tensor:
        """Buffer holding the weights in the activation dtype (bf16 shadow, refreshed, or the master itself)."""
        if act_dtype() == torch.float32:
            return st.flat
        sh = st.shadow()
        if refresh:
            ops.cast_f32_to_bf16(st.flat, sh)
        return sh


# ---------------------------------------------------------------------------------------------
# MultiHeadSelfAttention
# ---------------------------------------------------------------------------------------------
class _MHSAFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, module, drop, wq, wk, wv, bq, bk, bv, wo, bo):
        st = module._ensure_packed()
        dm = module._dims(x)
        cbuf = module._compute_buffer(st)
        c = LayerViews(st.layout, cbuf, "", dm.H, 0, False, with_ln=False, attn_prefix="")
        p = LayerViews(st.layout, st.flat, "", dm.H, 0, False, with_ln=False, attn_prefix="")
        alloc = Fn.default_alloc(x.device)
        xa = x.reshape(dm.rows, dm.H).to(act_dtype()).contiguous()
        am = torch.empty((dm.B, dm.heads, dm.T, dm.T), dtype=torch.float32, device=x.device) if module.save_attn_map else None
        y, saved = Fn.mhsa_fwd(xa, c, p, dm, alloc, residual=None, attn_map=am, drop=drop)
        if am is not None:
            module.attn_map = am
        ctx.saved = (saved, c, dm, st, x.dtype, x.shape, drop)
        return y.view(dm.B, dm.T, dm.H).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        saved, c, dm, st, xdt, xshape, drop = ctx.saved
        gbuf = torch.zeros(st.layout.total, dtype=torch.float32, device=dy.device)
        g = LayerViews(st.layout, gbuf, "", dm.H, 0, False, with_ln=False, attn_prefix="")
        dya = dy.reshape(dm.rows, dm.H).to(act_dtype()).contiguous()
        dx = Fn.mhsa_bwd(dya, saved, c, g, dm, Fn.default_alloc(dy.device), drop=drop)
        H = dm.H
        return (dx.view(xshape).to(xdt), None, None, g.wqkv[:H], g.wqkv[H:2 * H], g.wqkv[2 * H:], g.bqkv[:H], g.bqkv[H:2 * H], g.bqkv[2 * H:],
                g.wo, g.bo)


class MultiHeadSelfAttention(nn.Module, _FlatRoot, _DropState):
    """layers.py:68-103: three Linear(F,F) projections, softmax(QKᵀ/sqrt(F)), PV, out_project, dropout."""

    def __init__(self, features: int, head: int = 8, dropout: float = 0.0, save_attn_map: bool = False):
        super().__init__()
        self.head = head
        self.features = features
        self.sqrt_d = self.features ** 0.5
        self.Wq = nn.Linear(features, features)
        self.Wk = nn.Linear(features, features)
        self.Wv = nn.Linear(features, features)
        self.out_project = nn.Linear(features, features)
        self.dropout = nn.Dropout(dropout)
        self.save_attn_map = save_attn_map
        self._drop_seed, self._drop_calls = None, 0
        object.__setattr__(self, "_store", None)

    def _layout(self) -> FlatLayout:
        return FlatLayout(layer_entries("", self.features, 0, False, with_ln=False, attn_prefix=""))

    def _dims(self, x: torch.Tensor) -> Fn.Dims:
        B, T, F_ = x.shape
        dm = Fn.Dims(B=B, T=T, H=F_, heads=self.head, M=0, use_mlp=False)
        dm.check()
        return dm

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        drop = self._next_drop(self.dropout.p, self.training)  # layers.py:102
        return _MHSAFn.apply(x, self, drop, self.Wq.weight, self.Wk.weight, self.Wv.weight, self.Wq.bias, self.Wk.bias, self.Wv.bias,
                             self.out_project.weight, self.out_project.bias)


# ---------------------------------------------------------------------------------------------
# TransformerEncoder
# ---------------------------------------------------------------------------------------------
class _EncoderFn(torch.autograd.Function):
    """One autograd node per encoder block: fused kernels inside, module interface outside."""

    @staticmethod
    def forward(ctx, x, module, views, drop, *params):
        layout, pflat, cbuf, prefix = views
        dm = module._dims(x)
        p = LayerViews(layout, pflat, prefix, dm.H, dm.M, dm.use_mlp)
        c = LayerViews(layout, cbuf, prefix, dm.H, dm.M, dm.use_mlp)
        xa = x.reshape(dm.rows, dm.H)
        if xa.dtype != act_dtype():
            xa = xa.to(act_dtype())
        xa = xa.contiguous()
        am = torch.empty((dm.B, dm.heads, dm.T, dm.T), dtype=torch.float32, device=x.device) if module._save_attn_map else None
        y, saved = Fn.encoder_fwd(xa, c, p, dm, Fn.default_alloc(x.device), attn_map=am, drop=drop)
        if am is not None:
            module.attention.attn_map = am
        ctx.saved = (saved, c, p, dm, x.dtype, x.shape, module, drop)
        return y.view(dm.B, dm.T, dm.H).to(x.dtype)

    @staticmethod
    def backward(ctx, dy):
        saved, c, p, dm, xdt, xshape, module, drop = ctx.saved
        lay = module._own_layout()
        gbuf = torch.zeros(lay.total, dtype=torch.float32, device=dy.device)
        g = LayerViews(lay, gbuf, "", dm.H, dm.M, dm.use_mlp)
        dya = dy.reshape(dm.rows, dm.H)
        if dya.dtype != act_dtype():
            dya = dya.to(act_dtype())
        dx = Fn.encoder_bwd(dya.contiguous(), saved, c, p, g, dm, Fn.default_alloc(dy.device), drop=drop)
        H = dm.H
        grads = [g.ln1_w, g.ln1_b, g.wqkv[:H], g.wqkv[H:2 * H], g.wqkv[2 * H:], g.bqkv[:H], g.bqkv[H:2 * H], g.bqkv[2 * H:], g.wo, g.bo]
        if dm.use_mlp:
            grads += [g.ln2_w, g.ln2_b, g.w1, g.b1, g.w2, g.b2]
        return (dx.view(xshape).to(xdt), None, None, None, *grads)


class TransformerEncoder(nn.Module, _FlatRoot, _DropState):
    """layers.py:15-65: out = attention(la1(x)) + x ; out = mlp(la2(out)) + out  (Linear-GELU-Linear-GELU MLP)."""

    def __init__(self, features: int, mlp_hidden: int, head: int = 8, dropout: float = 0.0, use_mlp: bool = True,
                 save_attn_map: bool = False):
        super().__init__()
        self.la1 = nn.LayerNorm(features)
        self.attention = MultiHeadSelfAttention(features, head=head, dropout=dropout, save_attn_map=save_attn_map)
        self.la2 = nn.LayerNorm(features)
        if use_mlp:
            self.mlp = nn.Sequential(
                nn.Linear(features, mlp_hidden), nn.GELU(), nn.Dropout(dropout),
                nn.Linear(mlp_hidden, features), nn.GELU(), nn.Dropout(dropout))
        else:
            self.mlp = None
        self._save_attn_map = save_attn_map
        self._features, self._mlp_hidden, self._head, self._p_drop = features, mlp_hidden, head, dropout
        self._drop_seed, self._drop_calls = None, 0
        object.__setattr__(self, "_store", None)
        object.__setattr__(self, "_parent_views", None)  # set by a ViT that packed this block into its own buffer

    # -- storage ------------------------------------------------------------------------------
    def _own_layout(self) -> FlatLayout:
        return FlatLayout(layer_entries("", self._features, self._mlp_hidden, self.mlp is not None))

    _layout = _own_layout

    def _dims(self, x: torch.Tensor) -> Fn.Dims:
        B, T, F_ = x.shape
        if F_ != self._features:
            raise ValueError(f"expected {self._features} features, got {F_}")
        dm = Fn.Dims(B=B, T=T, H=F_, heads=self._head, M=self._mlp_hidden, use_mlp=self.mlp is not None)
        dm.check()
        return dm

    def _param_list(self):
        a = self.attention
        ps = [self.la1.weight, self.la1.bias, a.Wq.weight, a.Wk.weight, a.Wv.weight, a.Wq.bias, a.Wk.bias, a.Wv.bias,
              a.out_project.weight, a.out_project.bias]
        if self.mlp is not None:
            ps += [self.la2.weight, self.la2.bias, self.mlp[0].weight, self.mlp[0].bias, self.mlp[3].weight, self.mlp[3].bias]
        return ps

    # -- reference interface ------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        drop = self._next_drop(self._p_drop, self.training)  # one stream for the block's three nn.Dropout sites
        pv = self._parent_views
        if pv is not None and pv[4]():  # packed inside a ViT whose storage is still current
            views = pv[:4]
        else:
            st = self._ensure_packed()
            views = (st.layout, st.flat, self._compute_buffer(st), "")
        return _EncoderFn.apply(x, self, views, drop, *self._param_list())

    @property
    def save_attn_map(self):
        return self._save_attn_map

    @save_attn_map.setter
    def save_attn_map(self, value):
        self._save_attn_map = value
        self.attention.save_attn_map = value

    def get_attention_map(self):
        if self._save_attn_map:
            return self.attention.attn_map
        raise Exception("Attention map was not saved. Set save_attn_map=True when initializing the model.")
