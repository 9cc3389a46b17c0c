"""Drop-in ``ViT`` (reference: vit.py:19-89) whose forward/backward run on libvitb200's sm_100a kernels.

Constructor signature, defaults, sub-module names (``emb, cls_token, pos_emb, enc[i], fc``) and therefore the
``state_dict`` keys are the reference's, and sub-modules are created in the reference's order with the same
torch constructors, so the same seed gives the same initial weights.  ``patch`` is the NUMBER of patches per
side (vit.py:37).  ``in_c`` is accepted and ignored exactly as in the reference (vit.py:41 hard-codes 3).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import functional as Fn
from .layers import TransformerEncoder, _FlatRoot, act_dtype
from .params import FlatLayout, FlatStore, layer_entries


class _StemFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img, model, emb_w_c, emb_w, emb_b, pos, cls):
        img = img.contiguous().float()
        x0, words = Fn.stem_fwd(img, emb_w.detach(), emb_w_c, emb_b.detach(), None if cls is None else cls.detach().view(-1),
                                pos.detach().view(pos.shape[-2], pos.shape[-1]), model.patch, act_dtype(), Fn.default_alloc(img.device))
        ctx.saved = (img, model, cls is not None, words)
        B = img.shape[0]
        return x0.view(B, -1, x0.shape[-1])

    @staticmethod
    def backward(ctx, dx0):
        img, model, has_cls, words = ctx.saved
        H, K, T = model.hidden, model.emb.in_features, model.num_tokens
        dev = dx0.device
        g_w = torch.empty((H, K), dtype=torch.float32, device=dev)
        g_b = torch.empty((H,), dtype=torch.float32, device=dev)
        g_pos = torch.empty((T, H), dtype=torch.float32, device=dev)
        g_cls = torch.empty((H,), dtype=torch.float32, device=dev) if has_cls else None
        Fn.stem_bwd(img, words, dx0.reshape(-1, H).contiguous(), g_w, g_b, g_cls, g_pos, model.patch)
        return None, None, None, g_w, g_b, g_pos.view(1, T, H), (g_cls.view(1, 1, H) if has_cls else None)


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, model, fc_w_c, ln_w, ln_b, fc_w, fc_b):
        B, T, H = x.shape
        C = fc_w.shape[0]
        xa = x.reshape(B * T, H).contiguous()
        logits, saved = Fn.head_fwd(xa, ln_w.detach(), ln_b.detach(), fc_w_c, fc_b.detach(), B, T, H, C, model.is_cls_token,
                                    Fn.default_alloc(x.device))
        ctx.saved = (saved, ln_w.detach(), fc_w_c, (B, T, H, C), model.is_cls_token, x.dtype)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        saved, ln_w, fc_w_c, (B, T, H, C), is_cls, act = ctx.saved
        dev = dlogits.device
        g_ln_w = torch.empty((H,), dtype=torch.float32, device=dev)
        g_ln_b = torch.empty((H,), dtype=torch.float32, device=dev)
        g_fc_w = torch.empty((C, H), dtype=torch.float32, device=dev)
        g_fc_b = torch.empty((C,), dtype=torch.float32, device=dev)
        dx = Fn.head_bwd(dlogits.contiguous().float(), saved, ln_w, fc_w_c, g_ln_w, g_ln_b, g_fc_w, g_fc_b, B, T, H, C, is_cls, act,
                         Fn.default_alloc(dev))
        return dx.view(B, T, H), None, None, g_ln_w, g_ln_b, g_fc_w, g_fc_b


class ViT(nn.Module, _FlatRoot):
    def __init__(
        self,
        in_c: int = 3,
        num_classes: int = 10,
        img_size: int = 224,
        patch: int = 16,
        dropout: float = 0.0,
        num_layers: int = 12,
        hidden: int = 768,
        encoder_mlp: bool = True,
        mlp_hidden: int = 768 * 4,
        head: int = 8,
        is_cls_token: bool = True,
    ):
        super().__init__()
        self.patch = patch  # number of patches in one row (or column), vit.py:37
        self.is_cls_token = is_cls_token
        self.patch_size = img_size // self.patch
        assert self.patch_size * self.patch == img_size, "img_size must be divisible by patch"
        f = (img_size // self.patch) ** 2 * 3  # patch vector length, vit.py:41
        num_tokens = (self.patch ** 2) + 1 if self.is_cls_token else (self.patch ** 2)
        self.hidden, self.mlp_hidden, self.num_tokens, self.num_classes = hidden, mlp_hidden, num_tokens, num_classes
        self.img_size, self.num_layers, self.head, self.encoder_mlp, self.p_drop = img_size, num_layers, head, encoder_mlp, dropout

        # same construction order as the reference -> same RNG stream -> same initial weights for a given seed
        self.emb = nn.Linear(f, hidden)
        self.cls_token = nn.Parameter(torch.randn(1, 1, hidden)) if is_cls_token else None
        self.pos_emb = nn.Parameter(torch.randn(1, num_tokens, hidden))
        self.enc = nn.Sequential(*[
            TransformerEncoder(features=hidden, mlp_hidden=mlp_hidden, dropout=dropout, head=head, use_mlp=encoder_mlp)
            for _ in range(num_layers)])
        self.fc = nn.Sequential(nn.LayerNorm(hidden), nn.Linear(hidden, num_classes))
        object.__setattr__(self, "_store", None)

    # -- storage ------------------------------------------------------------------------------
    def _layout(self) -> FlatLayout:
        H, M, K, T, C = self.hidden, self.mlp_hidden, self.emb.in_features, self.num_tokens, self.num_classes
        e = []
        if self.is_cls_token:
            e.append(("cls_token", (1, 1, H), True))
        e += [("pos_emb", (1, T, H), True), ("emb.weight", (H, K), True), ("emb.bias", (H,), True)]
        for i in range(self.num_layers):
            e += layer_entries(f"enc.{i}.", H, M, self.encoder_mlp)
        e += [("fc.0.weight", (H,), True), ("fc.0.bias", (H,), True), ("fc.1.weight", (C, H), True), ("fc.1.bias", (C,), True)]
        return FlatLayout(e)

    def _after_pack(self, store: FlatStore) -> None:
        for i, blk in enumerate(self.enc):
            object.__setattr__(blk, "_parent_views", None)

    def bucket_bounds(self):
        """[(start, end)] element ranges of the flat buffer: stem, each encoder layer, head (gradient buckets)."""
        layout = self._layout()  # pure host arithmetic: usable without a device
        s = layout.slots
        bounds = []
        first_layer = s["enc.0.la1.weight"].off if self.num_layers > 0 else s["fc.0.weight"].off
        bounds.append((0, first_layer))
        for i in range(self.num_layers):
            a = s[f"enc.{i}.la1.weight"].off
            b = s[f"enc.{i + 1}.la1.weight"].off if i + 1 < self.num_layers else s["fc.0.weight"].off
            bounds.append((a, b))
        bounds.append((s["fc.0.weight"].off, layout.active_end))
        return bounds

    # -- reference interface ------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        st = self._ensure_packed()
        Fn.Dims(B=x.shape[0], T=self.num_tokens, H=self.hidden, heads=self.head, M=self.mlp_hidden, use_mlp=self.encoder_mlp).check()
        cbuf = self._compute_buffer(st)  # bf16 shadow refreshed from the fp32 master (one cast kernel)
        current = st.consistent
        for i, blk in enumerate(self.enc):
            object.__setattr__(blk, "_parent_views", (st.layout, st.flat, cbuf, f"enc.{i}.", current))
        emb_w_c = st.layout.view(cbuf, "emb.weight") if cbuf is not st.flat else None
        out = _StemFn.apply(x, self, emb_w_c, self.emb.weight, self.emb.bias, self.pos_emb, self.cls_token)  # vit.py:66-70
        out = self.enc(out)                                                                     # vit.py:71
        fc_w_c = st.layout.view(cbuf, "fc.1.weight")
        return _HeadFn.apply(out, self, fc_w_c, self.fc[0].weight, self.fc[0].bias, self.fc[1].weight, self.fc[1].bias)  # vit.py:72-76
