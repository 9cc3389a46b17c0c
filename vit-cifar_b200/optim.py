"""``FusedAdam``: torch.optim.Adam semantics (coupled L2 weight decay, as configured at network.py:71-77) in ONE
kernel launch over the model's flat parameter buffer, also refreshing the bf16 weight shadow."""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import ops


def adam_hyper(step: int, lr: float, beta1: float, beta2: float, eps: float, weight_decay: float, grad_scale: float = 1.0):
    """The 9 scalars of vitb_adam_multi, computed in double like torch.optim.Adam's single-tensor path."""
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    return [lr / bc1, math.sqrt(bc2), beta1, beta2, eps, weight_decay, grad_scale, 1.0 - beta1, 1.0 - beta2]


def sgd_hyper(lr: float, momentum: float, weight_decay: float, grad_scale: float = 1.0):
    """The scalars of vitb_sgd_multi in the slots of Adam's block: [0] lr, [2] momentum, [5] weight decay, [6] gradient scale."""
    return [lr, 0.0, momentum, 0.0, 0.0, weight_decay, grad_scale, 0.0, 0.0]


class FusedAdam:
    """Optimiser over a module packed by vit-cifar_b200 (``module._ensure_packed()``).

    Gradients are gathered from ``param.grad`` into the flat gradient buffer unless they already live there
    (the training engine writes them in place).  Parameters whose ``.grad`` is None are skipped, like torch.
    """

    def __init__(self, module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        self.module = module
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self._m: Optional[torch.Tensor] = None
        self._v: Optional[torch.Tensor] = None
        self._g: Optional[torch.Tensor] = None
        self._store = None

    def _state(self):
        st = self.module._ensure_packed()
        if self._store is not st:
            self._store = st
            self._m = torch.zeros_like(st.flat)
            self._v = torch.zeros_like(st.flat)
            self._g = torch.zeros_like(st.flat)
        return st

    def zero_grad(self, set_to_none: bool = True):
        for p in self.module.parameters():
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0):
        st = self._state()
        named = dict(self.module.named_parameters())
        self._g.zero_()
        for k, s in st.layout.slots.items():
            gr = named[k].grad
            if gr is not None:
                self._g[s.off:s.off + s.numel].copy_(gr.reshape(-1))
        self.step_count += 1
        h = adam_hyper(self.step_count, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, grad_scale)
        n = st.layout.active_end
        ops.adam(st.flat[:n], self._g[:n], self._m[:n], self._v[:n], st._shadow[:n] if st._shadow is not None else None, hyper_host=h)
