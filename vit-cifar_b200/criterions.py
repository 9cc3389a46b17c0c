"""Drop-in ``LabelSmoothingCrossEntropyLoss`` (reference: criterions.py:5-19) as one fused kernel.

Not ``F.cross_entropy(label_smoothing=)``: the off-target mass is s/(C-1) and the target gets exactly 1-s
(criterions.py:16-18).  Forward and the gradient (softmax - q)/B come out of the same pass.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


class _LSCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, smoothing):
        logits = pred.detach().float().contiguous()
        loss = torch.empty((), dtype=torch.float32, device=pred.device)
        dlogits = torch.empty_like(logits) if pred.requires_grad else None
        ops.ls_ce(logits, target.contiguous(), loss, dlogits, smoothing, 1.0)
        ctx.dlogits = dlogits
        ctx.in_dtype = pred.dtype
        return loss

    @staticmethod
    def backward(ctx, gout):
        if ctx.dlogits is None:
            return None, None, None
        return (ctx.dlogits * gout).to(ctx.in_dtype), None, None


class LabelSmoothingCrossEntropyLoss(nn.Module):
    def __init__(self, classes, smoothing=0.0, dim=-1):
        super().__init__()
        self.confidence = 1.0 - smoothing
        self.smoothing = smoothing
        self.cls = classes
        self.dim = dim

    def forward(self, pred, target):
        if pred.dim() != 2 or self.dim not in (-1, 1):
            raise ValueError("fused LS-CE expects (B, C) logits with the class dimension last")
        if pred.shape[1] != self.cls:
            raise ValueError(f"expected {self.cls} classes, got {pred.shape[1]}")
        return _LSCEFn.apply(pred, target, float(self.smoothing))
